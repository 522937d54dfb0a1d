// Round-2 groundwork for the weight-stationary recurrence: what does one stage exchange cost when a group
// of G co-resident CTAs shares a dialogue tile?  Measures, with cooperative launch on all SMs,
//   (1) a split arrive/wait barrier among G CTAs through one global counter (the sps kernels' scheme, per group),
//   (2) the same plus the exchange itself: every CTA writes its [slice x NB] fp32 block to global memory, and after
//       the barrier pulls the whole [K x NB] activation (all slices) into shared memory with one cp.async.bulk per slice,
//   (3) a hardware cluster barrier (cluster size 8, and 16 with the non-portable opt-in) with a DSMEM all-gather of the
//       same data.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/group_sync profiles/microbench/group_sync.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- (1)+(2): software group barrier through L2, optional exchange --------------------------------------------
__global__ void __launch_bounds__(256, 1) group_kernel(unsigned *bars, float *xbuf, int G, int iters, int slice_floats, int do_xchg) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long mbar;
    const int grp = blockIdx.x / G, rank = blockIdx.x % G, tid = threadIdx.x;
    unsigned *bar = bars + grp * 32;                       // one 128-byte line per group
    float *gx = xbuf + (size_t)grp * 2 * G * slice_floats;  // double-buffered exchange area of the group
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        float *buf = gx + (size_t)(it & 1) * G * slice_floats;
        if (do_xchg)                                       // my slice: coalesced float4 stores
            for (int i = tid; i < slice_floats / 4; i += blockDim.x)
                reinterpret_cast<float4 *>(buf + (size_t)rank * slice_floats)[i] = make_float4(it, rank, i, 0.f);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(bar, 1u);
            unsigned v, target = (unsigned)(it + 1) * G;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
            __threadfence();
            if (do_xchg) {                                 // pull all G slices with bulk copies
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(G * slice_floats * 4) : "memory");
                for (int r = 0; r < G; ++r)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(sm + (size_t)r * slice_floats)), "l"(buf + (size_t)r * slice_floats), "r"(slice_floats * 4),
                                 "r"(smem_u32(&mbar)) : "memory");
            }
        }
        __syncthreads();
        if (do_xchg) {
            unsigned ok = 0;
            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(smem_u32(&mbar)), "r"((unsigned)(it & 1)) : "memory");
        }
    }
    if (do_xchg && tid == 0 && sm[1] < -1.f) printf("x");
}

// ---- (3): hardware cluster barrier + DSMEM all-gather -----------------------------------------------------------
__global__ void __launch_bounds__(256, 1) cluster_kernel(int iters, int slice_floats, int do_xchg) {
    extern __shared__ __align__(128) float sm[];           // [2][CS][slice] : every CTA holds the gathered activation, double-buffered
    cg::cluster_group cl = cg::this_cluster();
    const int CS = cl.num_blocks(), rank = cl.block_rank(), tid = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        float *mine = sm + ((size_t)(it & 1) * CS + rank) * slice_floats;
        if (do_xchg) {
            // push my slice into every peer's buffer (remote stores, float4)
            for (int p = 0; p < CS; ++p) {
                float *dst = cl.map_shared_rank(mine, p);
                for (int i = tid; i < slice_floats / 4; i += blockDim.x) reinterpret_cast<float4 *>(dst)[i] = make_float4(it, rank, i, 0.f);
            }
        }
        cl.sync();
    }
    if (do_xchg && tid == 0 && sm[1] < -1.f) printf("x");
}

static float time_launch(void *fn, dim3 grid, dim3 block, void **args, size_t smem, bool coop, int cluster, void *zero = nullptr) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        if (zero) CK(cudaMemset(zero, 0, 64 * 128));      // barrier counters start from 0 in every repetition
        CK(cudaEventRecord(e0));
        if (cluster) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
            if (e != cudaSuccess) { printf("  cluster launch failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return -1.f; }
        } else if (coop) CK(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, 0));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int iters = 2000, NB = 96;                       // 96 dialogues per group (round-2 plan)
    unsigned *bars; float *xbuf;
    CK(cudaMalloc(&bars, 64 * 128)); CK(cudaMalloc(&xbuf, (size_t)256 << 20));
    CK(cudaFuncSetAttribute(group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    printf("SMs %d, %d iterations per measurement, NB = %d dialogues per group\n", sms, iters, NB);
    const int Gs[] = {8, 13, 16, 37, 148};
    for (int G : Gs) {
        const int groups = sms / G, grid = groups * G;
        for (int K : {0, 64, 208, 416}) {                  // activation width gathered per stage (0 = barrier only)
            int slice = K ? (K * NB + G - 1) / G : 0; slice = (slice + 3) & ~3;
            if ((size_t)G * slice * 4 > 190 * 1024) continue;
            CK(cudaMemset(bars, 0, 64 * 128));
            int g = G, it = iters, sf = slice, dx = K ? 1 : 0;
            void *args[] = {&bars, &xbuf, &g, &it, &sf, &dx};
            float ms = time_launch((void *)group_kernel, dim3(grid), dim3(256), args, (size_t)G * slice * 4 + 16, true, 0, bars);
            printf("L2 group barrier  G=%3d (%2d groups)  gather K=%3d (%6.1f KB per CTA per stage): %7.3f us per stage\n", G, groups, K,
                   G * slice * 4 / 1024.0, ms * 1e3 / iters);
        }
    }
    for (int CS : {8, 16}) {
        if (CS > 8) { cudaError_t e = cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); if (e != cudaSuccess) { printf("cluster %d not allowed\n", CS); continue; } }
        for (int K : {0, 64, 208}) {
            int slice = K ? (K * NB + CS - 1) / CS : 0; slice = (slice + 3) & ~3;
            size_t smem = (size_t)2 * CS * slice * 4 + 16;
            if (smem > 200 * 1024) continue;
            CK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            int it = iters, sf = slice, dx = K ? 1 : 0;
            void *args[] = {&it, &sf, &dx};
            int maxc = 0;
            { cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3((sms / CS) * CS); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
              cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
              cfg.attrs = at; cfg.numAttrs = 1; cudaOccupancyMaxActiveClusters(&maxc, (void *)cluster_kernel, &cfg); }
            const int nclusters = maxc < sms / CS ? maxc : sms / CS;
            if (nclusters < 1) { printf("cluster size %d: no resident clusters\n", CS); continue; }
            float ms = time_launch((void *)cluster_kernel, dim3(nclusters * CS), dim3(256), args, smem, false, CS);
            if (ms > 0) printf("cluster barrier   CS=%2d (%2d co-resident clusters of max %d)  DSMEM all-gather K=%3d: %7.3f us per stage\n", CS, nclusters, maxc, K, ms * 1e3 / iters);
        }
    }
    return 0;
}
