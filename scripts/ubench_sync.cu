// Micro-benchmark of the synchronisation primitives inside the GEMM main loop (bring-up tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I textmae_image_compression_b200/csrc scripts/ubench_sync.cu -o gpurun_out/ubench_sync
// Prints SM cycles per iteration of each primitive / loop shape.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace tmae;

constexpr int kIters = 2000;
#ifndef KSTAGES
#define KSTAGES 8
#endif
constexpr int kStages = KSTAGES;

__device__ __forceinline__ bool test_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.relaxed.cta.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool try_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void expect_tx_relaxed(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cta.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// mode 0..6: single-warp primitive loops;  mode 10..: two-warp pipelines
template <int mode>
__global__ void __launch_bounds__(64, 1) ubench(long long* out) {
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages], self;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&self, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(&tmem_slot, 32); tmem_relinquish(); }
    __syncthreads();
    const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty), self_a = smem_u32(&self);
    long long t0 = 0, t1 = 0;
    if (mode < 10) {
        if (warp == 0) {
            if (mode == 1 || mode == 5) {}  // barrier starts in phase 0: waiting on parity 1 returns immediately
            __syncwarp();
            t0 = clock64();
            uint32_t ph = 0;
            for (int i = 0; i < kIters; ++i) {
                switch (mode) {
                    case 0: mbar_try_wait_a(self_a, 1u); break;                                // completed phase
                    case 1: if (elect_one()) mbar_arrive(&self); __syncwarp(); mbar_wait_a(self_a, ph); ph ^= 1u; break;
                    case 2: if (elect_one()) { asm volatile("" ::: "memory"); } __syncwarp(); break;
                    case 3: fence_after(); break;
                    case 4: if (elect_one()) umma_commit_a(self_a); __syncwarp(); mbar_wait_a(self_a, ph); ph ^= 1u; break;
                    case 5: if (elect_one()) mbar_arrive_expect_tx_a(self_a, 0); __syncwarp(); mbar_wait_a(self_a, ph); ph ^= 1u; break;
                    case 6: if (lane == 0) mbar_arrive(&self); mbar_wait_a(self_a, ph); ph ^= 1u; break;   // lane-0 arrive, all wait
                }
            }
            t1 = clock64();
            if (lane == 0) out[0] = t1 - t0;
        }
    } else {
        // two-role pipeline over kStages: warp 0 = "producer" (wait empty, arrive full), warp 1 = "consumer" (wait full, free slot)
        //   10: warp-uniform loops, elect per iteration, consumer frees with tcgen05.commit
        //   11: same, consumer frees with mbarrier.arrive
        //   12: single-thread loops (lane 0 only), commit
        //   13: single-thread loops, arrive
        //   14: like 10 plus tcgen05.fence::after_thread_sync after the full wait
        if (mode >= 20 && mode < 30) {
            // one role alone: the waited barriers are always complete (fresh barrier, parity 1), the signalled barriers
            // never complete (huge count): intrinsic cost of wait -> signal per iteration
            __shared__ __align__(8) uint64_t sink[kStages];
            if (threadIdx.x == 0) { for (int i = 0; i < kStages; ++i) mbar_init(&sink[i], 1000000); fence_barrier_init(); }
            __syncthreads();
            const uint32_t sink_a = smem_u32(sink);
            if (mode == 29) {      // like 28 with relaxed test_wait / expect_tx
                const uint32_t my_a = warp == 0 ? empty_a : full_a;
                __syncthreads();
                t0 = clock64();
                int stage = 0;
                for (int i = 0; i < kIters; ++i) {
                    if (!test_wait_relaxed(my_a + 8u * stage, 1u)) __trap();
                    if (elect_one()) { if (warp == 0) expect_tx_relaxed(sink_a + 8u * stage, 0); else umma_commit_a(sink_a + 8u * ((stage + 4) % kStages)); }
                    __syncwarp();
                    if (++stage == kStages) stage = 0;
                }
                t1 = clock64();
                if (lane == 0) out[warp] = t1 - t0;
                __syncthreads();
                if (warp == 1) tmem_dealloc(tmem_slot, 32);
                return;
            }
            if (mode == 27 || mode == 28) {
                const uint32_t my_a = warp == 0 ? empty_a : full_a;
                __syncthreads();
                t0 = clock64();
                int stage = 0;
                for (int i = 0; i < kIters; ++i) {
                    if (!mbar_test_wait_a(my_a + 8u * stage, 1u)) __trap();
                    if (mode == 27) { if (elect_one()) { asm volatile("" ::: "memory"); } __syncwarp(); }
                    else { if (elect_one()) { if (warp == 0) mbar_arrive_expect_tx_a(sink_a + 8u * stage, 0); else umma_commit_a(sink_a + 8u * ((stage + 4) % kStages)); } __syncwarp(); }
                    if (++stage == kStages) stage = 0;
                }
                t1 = clock64();
                if (lane == 0) out[warp] = t1 - t0;
                __syncthreads();
                if (warp == 1) tmem_dealloc(tmem_slot, 32);
                return;
            }
            if (warp == 1) {
                t0 = clock64();
                int stage = 0;
                for (int i = 0; i < kIters; ++i) {
                    if (mode == 26) { if (!mbar_test_wait_a(full_a + 8u * stage, 1u)) __trap(); }
                    else if (mode != 25) mbar_wait_a(full_a + 8u * stage, 1u);
                    if (mode == 20) { if (elect_one()) mbar_arrive_expect_tx_a(sink_a + 8u * stage, 0); __syncwarp(); }
                    if (mode == 21 || mode == 25) { if (elect_one()) umma_commit_a(sink_a + 8u * stage); __syncwarp(); }
                    if (mode == 22 || mode == 26) { if (elect_one()) { asm volatile("" ::: "memory"); } __syncwarp(); }
                    if (mode == 23) { if (lane == 0) mbar_arrive(&sink[stage]); }
                    if (mode == 24) { if (elect_one()) mbar_arrive(&sink[stage]); __syncwarp(); }
                    if (++stage == kStages) stage = 0;
                }
                t1 = clock64();
                if (lane == 0) out[1] = t1 - t0;
            }
            __syncthreads();
            if (warp == 1) tmem_dealloc(tmem_slot, 32);
            return;
        }
        if (mode == 31 || mode == 32) {
            __syncthreads();
            t0 = clock64();
            int stage = 0;
            uint32_t phase = 0;
            long long nspin = 0;
            for (int i = 0; i < kIters; ++i) {
                const uint32_t wa = (warp == 0 ? empty_a : full_a) + 8u * stage;
                const uint32_t wp = warp == 0 ? phase ^ 1u : phase;
                if (mode == 31) { while (!mbar_try_wait_a(wa, wp)) { ++nspin; } }
                else { uint32_t spins = 0; while (!mbar_try_wait_a(wa, wp)) { ++nspin; if (++spins > 100000000u) asm volatile("trap;"); } }
                if (elect_one()) { if (warp == 0) mbar_arrive_expect_tx_a(full_a + 8u * stage, 0); else umma_commit_a(empty_a + 8u * stage); }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            t1 = clock64();
            if (lane == 0) { out[warp] = t1 - t0; out[2 + warp] = nspin; out[4 + warp] = t0; out[6 + warp] = t1; }
            __syncthreads();
            if (warp == 1) tmem_dealloc(tmem_slot, 32);
            return;
        }
        if (mode == 30) {
            __syncthreads();
            t0 = clock64();
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < kIters; ++i) {
                if (warp == 0) {
                    while (!try_wait_relaxed(empty_a + 8u * stage, phase ^ 1u)) {}
                    if (elect_one()) expect_tx_relaxed(full_a + 8u * stage, 0);
                    __syncwarp();
                } else {
                    while (!try_wait_relaxed(full_a + 8u * stage, phase)) {}
                    if (elect_one()) umma_commit_a(empty_a + 8u * stage);
                    __syncwarp();
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            t1 = clock64();
            if (lane == 0) out[warp] = t1 - t0;
            __syncthreads();
            if (warp == 1) tmem_dealloc(tmem_slot, 32);
            return;
        }
        if (mode == 17) {
            __syncthreads();
            t0 = clock64();
            int stage = 0, pstage = 0;
            uint32_t phase = 0, pphase = 0;
            const uint32_t my_a = warp == 0 ? empty_a : full_a, other_a = warp == 0 ? full_a : empty_a;
            const uint32_t inv = warp == 0 ? 1u : 0u;
            auto probe = [&]() { const bool r = mbar_test_wait_a(my_a + 8u * pstage, pphase ^ inv); if (++pstage == kStages) { pstage = 0; pphase ^= 1u; } return r; };
            bool r0 = probe(), r1 = probe();
            for (int i = 0; i < kIters; ++i) {
                const bool r2 = probe();
                if (!r0) mbar_wait_a(my_a + 8u * stage, phase ^ inv);
                if (elect_one()) { if (warp == 0) mbar_arrive_expect_tx_a(other_a + 8u * stage, 0); else umma_commit_a(other_a + 8u * stage); }
                __syncwarp();
                r0 = r1; r1 = r2;
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            t1 = clock64();
            if (lane == 0) out[warp] = t1 - t0;
            __syncthreads();
            if (warp == 1) tmem_dealloc(tmem_slot, 32);
            return;
        }
        if (mode == 15 || mode == 16) {
            // software-pipelined peek: the try_wait of the NEXT stage is issued before this stage's arrive / commit, its
            // predicate is consumed one iteration later (mode 16: the arrive comes first, then the peek)
            __syncthreads();
            t0 = clock64();
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t my_a = warp == 0 ? empty_a : full_a, other_a = warp == 0 ? full_a : empty_a;
            const uint32_t inv = warp == 0 ? 1u : 0u;
            bool ready = mbar_try_wait_a(my_a, phase ^ inv);
            for (int i = 0; i < kIters; ++i) {
                if (!ready) mbar_wait_a(my_a + 8u * stage, phase ^ inv);
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == kStages) { nstage = 0; nphase ^= 1u; }
                if (mode == 15) ready = mbar_try_wait_a(my_a + 8u * nstage, nphase ^ inv);
                if (elect_one()) { if (warp == 0) mbar_arrive_expect_tx_a(other_a + 8u * stage, 0); else umma_commit_a(other_a + 8u * stage); }
                __syncwarp();
                if (mode == 16) ready = mbar_try_wait_a(my_a + 8u * nstage, nphase ^ inv);
                stage = nstage; phase = nphase;
            }
            t1 = clock64();
            if (lane == 0) out[warp] = t1 - t0;
            __syncthreads();
            if (warp == 1) tmem_dealloc(tmem_slot, 32);
            return;
        }
        const bool single = mode == 12 || mode == 13;
        const bool use_commit = mode == 10 || mode == 12 || mode == 14;
        __syncthreads();
        t0 = clock64();
        if (!single || lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < kIters; ++i) {
                if (warp == 0) {
                    mbar_wait_a(empty_a + 8u * stage, phase ^ 1u);
                    if (single) mbar_arrive_expect_tx_a(full_a + 8u * stage, 0);
                    else { if (elect_one()) mbar_arrive_expect_tx_a(full_a + 8u * stage, 0); __syncwarp(); }
                } else {
                    mbar_wait_a(full_a + 8u * stage, phase);
                    if (mode == 14) fence_after();
                    if (single) { if (use_commit) umma_commit_a(empty_a + 8u * stage); else mbar_arrive(&empty[stage]); }
                    else { if (elect_one()) { if (use_commit) umma_commit_a(empty_a + 8u * stage); else mbar_arrive(&empty[stage]); } __syncwarp(); }
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
        t1 = clock64();
        if (lane == 0) out[warp] = t1 - t0;
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_slot, 32);
}

// ---- TMA load cost: one thread issues `batch` loads of a 64 x `rows` bf16 box (SWIZZLE_128B) per mbarrier phase ----
#include <cuda.h>
__global__ void __launch_bounds__(32, 1) tma_batch(const __grid_constant__ CUtensorMap map, int batch, int box_bytes, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    const uint32_t bar_a = smem_u32(&bar), base = smem_u32(sm);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx_a(bar_a, (uint32_t)(batch * box_bytes));
            for (int j = 0; j < batch; ++j) tma_load_2d_a(base + (uint32_t)(j * box_bytes), &map, bar_a, 0, ((i * batch + j) * 128) % 4096);
        }
        __syncwarp();
        mbar_wait_a(bar_a, ph);
        ph ^= 1u;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void tma_bench(long long* d) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    void* buf; cudaMalloc(&buf, 8192 * 64 * 2); cudaMemset(buf, 0, 8192 * 64 * 2);
    cudaFuncSetAttribute(tma_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rows : {128, 64, 32}) {
        CUtensorMap m; cuuint64_t gd[2] = {64, 8192}; cuuint64_t gs[1] = {128}; cuuint32_t bx[2] = {64, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
        ((PFN_enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        for (int batch : {1, 2, 4, 8}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) tma_batch<<<1, 32, 8 * 16384 + 1024>>>(m, batch, rows * 128, 500, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("tma box 64x%-3d (%5d B) x batch %d: %8.1f clk per phase, %7.1f clk per load  %s\n", rows, rows * 128, batch, (double)h / 500, (double)h / 500 / batch,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    tma_bench(d);
    const int modes[] = {0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 14, 15, 16, 20, 21, 22, 23, 24, 25, 26, 17, 27, 28, 29, 30, 31, 32};
    const char* names[] = {"try_wait (completed phase)", "elect arrive + wait (self)", "elect + syncwarp", "tcgen05.fence::after",
                           "elect tcgen05.commit + wait (self)", "elect arrive.expect_tx(0) + wait (self)", "lane0 arrive + wait (self)",
                           "pipeline: warp-uniform, commit", "pipeline: warp-uniform, arrive", "pipeline: single thread, commit",
                           "pipeline: single thread, arrive", "pipeline: warp-uniform, commit + fence",
                           "pipeline: peek next before arrive/commit", "pipeline: peek next after arrive/commit",
                           "alone: wait(ready) + elect expect_tx", "alone: wait(ready) + elect commit", "alone: wait(ready) + elect nothing",
                           "alone: wait(ready) + lane0 arrive", "alone: wait(ready) + elect arrive", "alone: elect commit only", "alone: test_wait(ready) + elect nothing",
                           "pipeline: test_wait probe 2 ahead", "both warps alone: test_wait + elect nothing", "both warps alone: test_wait + signal sink", "both warps alone: RELAXED test_wait + signal", "pipeline: RELAXED try_wait + expect_tx / commit", "pipeline: plain spin (acquire)", "pipeline: counted spin + trap (acquire)"};
    for (int m = 0; m < (int)(sizeof(modes) / sizeof(int)); ++m) {
        long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemset(d, 0, 64);
        for (int rep = 0; rep < 2; ++rep) {
            switch (modes[m]) {
                case 0: ubench<0><<<1, 64>>>(d); break;   case 1: ubench<1><<<1, 64>>>(d); break;
                case 2: ubench<2><<<1, 64>>>(d); break;   case 3: ubench<3><<<1, 64>>>(d); break;
                case 4: ubench<4><<<1, 64>>>(d); break;   case 5: ubench<5><<<1, 64>>>(d); break;
                case 6: ubench<6><<<1, 64>>>(d); break;   case 10: ubench<10><<<1, 64>>>(d); break;
                case 11: ubench<11><<<1, 64>>>(d); break; case 12: ubench<12><<<1, 64>>>(d); break;
                case 13: ubench<13><<<1, 64>>>(d); break; case 14: ubench<14><<<1, 64>>>(d); break;
                case 15: ubench<15><<<1, 64>>>(d); break; case 16: ubench<16><<<1, 64>>>(d); break;
                case 20: ubench<20><<<1, 64>>>(d); break; case 21: ubench<21><<<1, 64>>>(d); break;
                case 22: ubench<22><<<1, 64>>>(d); break; case 23: ubench<23><<<1, 64>>>(d); break;
                case 24: ubench<24><<<1, 64>>>(d); break; case 25: ubench<25><<<1, 64>>>(d); break;
                case 26: ubench<26><<<1, 64>>>(d); break; case 17: ubench<17><<<1, 64>>>(d); break;
                case 27: ubench<27><<<1, 64>>>(d); break; case 28: ubench<28><<<1, 64>>>(d); break;
                case 29: ubench<29><<<1, 64>>>(d); break; case 30: ubench<30><<<1, 64>>>(d); break;
                case 31: ubench<31><<<1, 64>>>(d); break; case 32: ubench<32><<<1, 64>>>(d); break;
            }
        }
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        if (modes[m] == 31 || modes[m] == 32) printf("   raw: dt0 %lld dt1 %lld spins %lld %lld  t0 %lld %lld t1 %lld %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
        printf("mode %2d %-44s %8.1f clk/iter (role0)  %8.1f clk/iter (role1)  %s\n", modes[m], names[m], (double)h[0] / kIters,
               (double)h[1] / kIters, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
