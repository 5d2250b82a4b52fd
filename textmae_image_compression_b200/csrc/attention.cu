// Fused softmax attention over the visible tokens of one image and one head (timm 0.4.5 Attention.forward as
// used by the encoder blocks, MCM.py:313-322, 629-630):  softmax((q k^T) * hd^-0.5) v, head_dim = 64, T <= ~300.
// One CTA per (image, head); K and V of all T keys stay in shared memory, each warp owns 16-query tiles and runs
// an online-softmax loop over 16-key chunks; S and P never leave registers.  bf16 tensor-core math
// (mma.sync m16n8k16, fp32 accumulate): this op is 1.7-3 % of the path's FLOPs, so it uses the legacy
// warp-level MMA; the GEMM / conv engine (gemm_tc.cu) is where tcgen05 is spent.
#include <math.h>

#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace tmae {

namespace {

constexpr int HD = 64;
constexpr int KPITCH = HD + 8;        // bf16 elements per K row in smem: 36 words -> conflict-free B-fragment loads
constexpr int kAttnWarps = 5;          // T = 65 (K = 64 kept patches + cls) is 5 query tiles of 16: one per warp, one round
constexpr int kAttnThreads = 32 * kAttnWarps;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// qkv: bf16 [N*T, 3C] with columns [3][H][64] (timm reshape (B,T,3,H,hd)); out: bf16 [N*T, C] columns [H][64].
// K and V of all keys are staged ROW-major (pitch 72 elements: the 8 rows of an ldmatrix tile fall in distinct bank
// groups) together with each warp's 16-row Q tile, all in one round of independent 16-byte global loads; every MMA
// fragment then comes from ldmatrix (Q, K plain; V .trans): no transposed copy of V, no scalar shared-memory traffic
// and no Q registers held across the key loop (<= 64 registers -> 6 CTAs / 30 warps per SM, 888 slots >= the 768
// (image, head) CTAs of batch 64: one wave, one query tile per warp).
__global__ void __launch_bounds__(kAttnThreads, 6)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int Tp, int H, int C,
                 float scale_log2e) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) uint8_t smem_attn[];
    __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn);            // [Tp][KPITCH]
    __nv_bfloat16* sV = sK + (size_t)Tp * KPITCH;                                // [Tp][KPITCH]
    __nv_bfloat16* sQ = sV + (size_t)Tp * KPITCH;                                // [kAttnWarps][16][KPITCH]
    const int n = blockIdx.x / H, h = blockIdx.x - n * H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const size_t ld = (size_t)3 * C;
    const __nv_bfloat16* base = qkv + (size_t)n * T * ld + (size_t)h * HD;
    __nv_bfloat16* myQ = sQ + (size_t)warp * 16 * KPITCH;

    // this warp's Q tile: 16 rows x 64 dims = 128 16-byte chunks, 4 per lane (rows >= T are zero)
    auto stage_q = [&](int q0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = i * 32 + lane, r = e >> 3, c8 = (e & 7) * 8;
            uint4 qv = make_uint4(0, 0, 0, 0);
            if (q0 + r < T) qv = *reinterpret_cast<const uint4*>(base + (size_t)(q0 + r) * ld + c8);
            *reinterpret_cast<uint4*>(myQ + r * KPITCH + c8) = qv;
        }
    };
    const int q_tiles = Tp / 16;
    if (warp < q_tiles) stage_q(warp * 16);
    // ---- stage K and V for all keys (16-byte chunks, coalesced: 8 threads per key row); zero the padding ----
#pragma unroll 4
    for (int e = tid; e < Tp * (HD / 8); e += kAttnThreads) {
        const int key = e / (HD / 8), c8 = (e - key * (HD / 8)) * 8;
        uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (key < T) {
            kv = *reinterpret_cast<const uint4*>(base + (size_t)key * ld + C + c8);
            vv = *reinterpret_cast<const uint4*>(base + (size_t)key * ld + 2 * C + c8);
        }
        *reinterpret_cast<uint4*>(sK + (size_t)key * KPITCH + c8) = kv;
        *reinterpret_cast<uint4*>(sV + (size_t)key * KPITCH + c8) = vv;
    }
    __syncthreads();

    const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV), sQ_a = smem_u32(myQ);
    // per-lane ldmatrix row addresses (bytes) relative to the first key of a chunk
    //   K (plain):  matrix j = lane >> 3 covers d columns [8j, 8j+8) of keys (lane & 7)           (+32 columns for the 2nd load)
    //   V (.trans): matrices (keys 0-7, d0) (keys 8-15, d0) (keys 0-7, d0+8) (keys 8-15, d0+8)
    //   Q (plain):  matrices (rows 0-7, k0) (rows 8-15, k0) (rows 0-7, k0+8) (rows 8-15, k0+8) = a0..a3 of one k-step
    const uint32_t k_lane = (uint32_t)(((lane & 7) * KPITCH + (lane >> 3) * 8) * 2);
    const uint32_t v_lane = (uint32_t)((((lane & 7) + ((lane >> 3) & 1) * 8) * KPITCH + (lane >> 4) * 8) * 2);

    for (int qt = warp; qt < q_tiles; qt += kAttnWarps) {
        const int q0 = qt * 16;
        if (qt != warp) {                          // later rounds (T > 80): restage this warp's Q slot
            __syncwarp();
            stage_q(q0);
            __syncwarp();
        }
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;     // rows g and g+8

        for (int k0 = 0; k0 < Tp; k0 += 16) {
            // S chunk = Q (16x64) * K[k0..k0+16)^T : two n-tiles of 8 keys
            float sacc[2][4];
            sacc[0][0] = sacc[0][1] = sacc[0][2] = sacc[0][3] = 0.f;
            sacc[1][0] = sacc[1][1] = sacc[1][2] = sacc[1][3] = 0.f;
            const uint32_t ka = sK_a + (uint32_t)(k0 * KPITCH * 2) + k_lane;
#pragma unroll
            for (int hk = 0; hk < 2; ++hk) {       // d 0..31 (k-steps 0, 1), then d 32..63 (k-steps 2, 3)
                uint32_t kb0[4], kb1[4], qa[4];
                ldmatrix_x4(kb0, ka + (uint32_t)(hk * 64));                               // keys k0 .. k0+7
                ldmatrix_x4(kb1, ka + (uint32_t)(8 * KPITCH * 2 + hk * 64));              // keys k0+8 .. k0+15
                ldmatrix_x4(qa, sQ_a + v_lane + (uint32_t)(hk * 64));                     // k-step 2 hk
                mma_bf16_16816(sacc[0], qa, kb0[0], kb0[1]);
                mma_bf16_16816(sacc[1], qa, kb1[0], kb1[1]);
                ldmatrix_x4(qa, sQ_a + v_lane + (uint32_t)(hk * 64 + 32));                // k-step 2 hk + 1
                mma_bf16_16816(sacc[0], qa, kb0[2], kb0[3]);
                mma_bf16_16816(sacc[1], qa, kb1[2], kb1[3]);
            }
            // scale (log2 domain), mask padded keys
            float cmax0 = -INFINITY, cmax1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int key = k0 + nt * 8 + 2 * t;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool ok = (key + (i & 1)) < T;
                    sacc[nt][i] = ok ? sacc[nt][i] * scale_log2e : -INFINITY;
                }
                cmax0 = fmaxf(cmax0, fmaxf(sacc[nt][0], sacc[nt][1]));
                cmax1 = fmaxf(cmax1, fmaxf(sacc[nt][2], sacc[nt][3]));
            }
            cmax0 = fmaxf(cmax0, __shfl_xor_sync(0xffffffffu, cmax0, 1));
            cmax0 = fmaxf(cmax0, __shfl_xor_sync(0xffffffffu, cmax0, 2));
            cmax1 = fmaxf(cmax1, __shfl_xor_sync(0xffffffffu, cmax1, 1));
            cmax1 = fmaxf(cmax1, __shfl_xor_sync(0xffffffffu, cmax1, 2));
            const float nm0 = fmaxf(m0, cmax0), nm1 = fmaxf(m1, cmax1);   // finite: key 0 of chunk 0 is always valid
            const float corr0 = exp2f(m0 - nm0), corr1 = exp2f(m1 - nm1);
            m0 = nm0; m1 = nm1;
            float rs0 = 0.f, rs1 = 0.f;
            uint32_t pa[4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const float p0 = exp2f(sacc[nt][0] - nm0), p1 = exp2f(sacc[nt][1] - nm0);
                const float p2 = exp2f(sacc[nt][2] - nm1), p3 = exp2f(sacc[nt][3] - nm1);
                rs0 += p0 + p1;
                rs1 += p2 + p3;
                pa[nt * 2 + 0] = pack_bf16x2(p0, p1);      // rows g   : a0 (keys 0-7) / a2 (keys 8-15)
                pa[nt * 2 + 1] = pack_bf16x2(p2, p3);      // rows g+8 : a1 / a3
            }
            l0 = l0 * corr0 + rs0;
            l1 = l1 * corr1 + rs1;
#pragma unroll
            for (int dt = 0; dt < 8; ++dt) {
                o[dt][0] *= corr0; o[dt][1] *= corr0; o[dt][2] *= corr1; o[dt][3] *= corr1;
            }
            // O += P (16 x 16 keys) * V[k0..k0+16) (16 keys x 64): one transposing ldmatrix per pair of 8-wide d tiles
            const uint32_t va = sV_a + (uint32_t)(k0 * KPITCH * 2) + v_lane;
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, va + (uint32_t)(dp * 32));
                mma_bf16_16816(o[2 * dp], pa, vb[0], vb[1]);
                mma_bf16_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
            }
        }
        // row sums live per quad: reduce across the 4 lanes that share a row
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.f / l0, inv1 = 1.f / l1;
        const int r0 = q0 + g, r1 = q0 + g + 8;
        __nv_bfloat16* obase = out + (size_t)n * T * C + (size_t)h * HD;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            const int c = dt * 8 + 2 * t;
            if (r0 < T) *reinterpret_cast<uint32_t*>(obase + (size_t)r0 * C + c) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
            if (r1 < T) *reinterpret_cast<uint32_t*>(obase + (size_t)r1 * C + c) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
        }
    }
}

inline int attn_tp(int T) { return (T + 15) / 16 * 16; }
inline size_t attn_smem(int T) {
    const int Tp = attn_tp(T);
    return ((size_t)2 * Tp + (size_t)kAttnWarps * 16) * KPITCH * sizeof(__nv_bfloat16);
}

// ---------------------------------------------------------------------------------------------------------
// Precise mode (TMAE_FLAG_PRECISE_ALL): fp32 softmax attention on CUDA cores.  q, k, v arrive as split-bf16 plane pairs
// (hi + lo = 16 mantissa bits) from the precise QKV GEMM; scores, softmax (expf, like the reference's fp32 softmax) and
// the PV sum run in fp32; the result leaves as a plane pair for the precise proj GEMM.  One CTA per (image, head):
// K (pitch 65, conflict-free for lane = key) and V (lane = dim) of all keys in shared memory, one warp per query row.
// This op is 0.8-3 % of the path's FLOPs; precise mode trades throughput for reference-exact symbols.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAttnF32Warps = 8;
constexpr int KP32 = HD + 1;
__global__ void __launch_bounds__(32 * kAttnF32Warps)
attention_f32_kernel(const __nv_bfloat16* __restrict__ qkv, long long qkv_lo, __nv_bfloat16* __restrict__ out, long long out_lo,
                     int T, int H, int C, float scale) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) uint8_t smem_attn[];
    float* sK = reinterpret_cast<float*>(smem_attn);                 // [T][65]
    float* sV = sK + (size_t)T * KP32;                               // [T][64]
    float* sQ = sV + (size_t)T * HD;                                 // [warps][64]
    float* sP = sQ + kAttnF32Warps * HD;                             // [warps][T]
    const int n = blockIdx.x / H, h = blockIdx.x - n * H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t ld = (size_t)3 * C;
    const __nv_bfloat16* base = qkv + (size_t)n * T * ld + (size_t)h * HD;
    auto ldf = [&](const __nv_bfloat16* p) { return load_bf16_planes(p, qkv_lo); };
    for (int e = tid; e < T * HD; e += blockDim.x) {
        const int key = e / HD, d = e - key * HD;
        sK[key * KP32 + d] = ldf(base + (size_t)key * ld + C + d);
        sV[key * HD + d] = ldf(base + (size_t)key * ld + 2 * C + d);
    }
    __syncthreads();
    float* myQ = sQ + warp * HD;
    float* myP = sP + (size_t)warp * T;
    for (int q = warp; q < T; q += kAttnF32Warps) {
        myQ[lane] = ldf(base + (size_t)q * ld + lane);
        myQ[lane + 32] = ldf(base + (size_t)q * ld + lane + 32);
        __syncwarp();
        float mx = -INFINITY;
        for (int j = lane; j < T; j += 32) {
            const float* kr = sK + j * KP32;
            float acc = 0.f;
#pragma unroll 16
            for (int d = 0; d < HD; ++d) acc = fmaf(myQ[d], kr[d], acc);
            acc *= scale;
            myP[j] = acc;
            mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < T; j += 32) {
            const float pj = expf(myP[j] - mx);
            myP[j] = pj;
            sum += pj;
        }
        sum = warp_sum(sum);
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < T; ++j) {
            const float pj = myP[j];
            o0 = fmaf(pj, sV[j * HD + lane], o0);
            o1 = fmaf(pj, sV[j * HD + lane + 32], o1);
        }
        const float inv = 1.f / sum;
        o0 *= inv; o1 *= inv;
        __nv_bfloat16* dst = out + ((size_t)n * T + q) * C + (size_t)h * HD;
        store_bf16_planes(dst + lane, out_lo, o0);
        store_bf16_planes(dst + lane + 32, out_lo, o1);
        __syncwarp();
    }
}
inline size_t attn_f32_smem(int T) {
    return ((size_t)T * KP32 + (size_t)T * HD + (size_t)kAttnF32Warps * HD + (size_t)kAttnF32Warps * T) * sizeof(float);
}

// ---------------------------------------------------------------------------------------------------------
// tcgen05 attention (bf16 mode, T <= 192): one CTA per (image, head).  Q, K, V tiles arrive by TMA straight out of the
// qkv matrix; S = Q K^T and O = P V are tcgen05.mma with the accumulators in tensor memory; softmax runs thread-per-row on
// the S row read back with tcgen05.ld (no shuffles), P goes to shared memory as the K-major 128B-swizzled A operand of the
// second MMA, V is transposed in shared memory into the K-major B operand ([64 dims][keys]) while the first MMA runs.
//   smem: Q [128 x 64] | K [Tp x 64] (both as TMA wrote them: SWIZZLE_128B, K-major) | V raw [Tp x 64] (no swizzle) |
//         V^T [64 x keys] in 64-key atoms | P [128 x keys] in 64-key atoms.   TMEM: S at column 0 (Tp <= 192), O at 192.
// Rows >= T of a tile belong to the next image (or are zero-filled at the end of the tensor): as keys they are masked in the
// softmax, as queries they are computed and not stored.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAttnTcThreads = 128;
constexpr int kAttnTcMaxTp = 192;
constexpr uint32_t kAttnOCol = 192;

__global__ void __launch_bounds__(kAttnTcThreads)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, __nv_bfloat16* __restrict__ out, int T, int Tp, int H, int C,
                    float scale_log2e) {
    pdl_launch_dependents();
    extern __shared__ __align__(1024) uint8_t smem_tc[];
    const int atoms = (Tp + 63) >> 6;
    const uint32_t kbytes = (uint32_t)Tp * 128u;
    const uint32_t kreg = (kbytes + 1023u) & ~1023u;
    uint8_t* sQ = smem_tc;                               // 16 KB
    uint8_t* sK = sQ + 16384;
    uint8_t* sVr = sK + kreg;
    uint8_t* sVt = sVr + kreg;                           // atoms x 8 KB
    uint8_t* sP = sVt + (size_t)atoms * 8192;            // atoms x 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (size_t)atoms * 16384);     // load, s, o
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    const int n = blockIdx.x / H, h = blockIdx.x - n * H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
        fence_barrier_init();
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    }
    if (warp == 0) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    const uint32_t q_a = smem_u32(sQ), k_a = smem_u32(sK), vr_a = smem_u32(sVr), vt_a = smem_u32(sVt), p_a = smem_u32(sP);
    const uint32_t bar_load = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_o = smem_u32(&bars[2]);
    const int row0 = n * T;
    const int q_tiles = (T + 127) >> 7;
    for (int qt = 0; qt < q_tiles; ++qt) {
        const uint32_t ph = (uint32_t)qt & 1u;
        const int q0 = qt * 128;
        if (tid == 0) {
            mbar_arrive_expect_tx_a(bar_load, 16384u + (qt == 0 ? 2u * kbytes : 0u));
            tma_load_2d_a(q_a, &map_q, bar_load, h * HD, row0 + q0);
            if (qt == 0) {
                tma_load_2d_a(k_a, &map_k, bar_load, C + h * HD, row0);
                tma_load_2d_a(vr_a, &map_v, bar_load, 2 * C + h * HD, row0);
            }
        }
        mbar_wait_a(bar_load, ph);
        if (tid == 0) {                                   // S = Q K^T: M = 128, N = Tp, K = 64 in four k-steps
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16_f32(128, Tp);
            const uint64_t dq = umma_smem_desc_sw128(q_a), dk = umma_smem_desc_sw128(k_a);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
            umma_commit_a(bar_s);
        }
        if (qt == 0) {
            // V^T while the MMA runs: thread -> (dim d, every second 8-key chunk); 8 strided 2-byte reads, one 16-byte store
            const int d = tid & 63;
            for (int j0 = (tid >> 6) * 8; j0 < Tp; j0 += 16) {
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint16_t lo = *reinterpret_cast<const uint16_t*>(sVr + (size_t)(j0 + 2 * i) * 128 + d * 2);
                    const uint16_t hi = *reinterpret_cast<const uint16_t*>(sVr + (size_t)(j0 + 2 * i + 1) * 128 + d * 2);
                    pk[i] = (uint32_t)lo | ((uint32_t)hi << 16);
                }
                const uint32_t dst = vt_a + (uint32_t)(j0 >> 6) * 8192u + (uint32_t)d * 128u + ((((uint32_t)(j0 & 63) >> 3) ^ (uint32_t)(d & 7)) << 4);
                sts128(dst, pk[0], pk[1], pk[2], pk[3]);
            }
        }
        // ---- softmax of this thread's query row (TMEM lane = row) ----
        mbar_wait_a(bar_s, ph);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        float mx = -INFINITY;
        for (int c0 = 0; c0 < Tp; c0 += 16) {
            uint32_t sv[16];
            tmem_ld_32x32b_x16(lane_base + (uint32_t)c0, sv);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c0 + i < T) mx = fmaxf(mx, __uint_as_float(sv[i]));
        }
        const float mxs = mx * scale_log2e;
        float sum = 0.f;
        const uint32_t prow = p_a + (uint32_t)row * 128u;
        for (int c0 = 0; c0 < Tp; c0 += 16) {
            uint32_t sv[16];
            tmem_ld_32x32b_x16(lane_base + (uint32_t)c0, sv);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float p0 = (c0 + 2 * i < T) ? exp2f(fmaf(__uint_as_float(sv[2 * i]), scale_log2e, -mxs)) : 0.f;
                const float p1 = (c0 + 2 * i + 1 < T) ? exp2f(fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2e, -mxs)) : 0.f;
                sum += p0 + p1;
                pk[i] = pack_bf16x2(p0, p1);
            }
            const uint32_t base = prow + (uint32_t)(c0 >> 6) * 16384u;
            const uint32_t ch = (uint32_t)(c0 & 63) >> 3;
            sts128(base + (((ch) ^ (uint32_t)(row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
            sts128(base + (((ch + 1u) ^ (uint32_t)(row & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async_smem();                         // P (and V^T) written through the generic proxy -> visible to the MMA
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {                                   // O = P V: M = 128, N = 64, K = Tp in Tp / 16 k-steps
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16_f32(128, HD);
            const uint64_t dp = umma_smem_desc_sw128(p_a), dv = umma_smem_desc_sw128(vt_a);
            const int ksteps = Tp >> 4;
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t a_desc = dp + (uint64_t)((uint32_t)(ks >> 2) * (16384u >> 4)) + (uint64_t)(2 * (ks & 3));
                const uint64_t b_desc = dv + (uint64_t)((uint32_t)(ks >> 2) * (8192u >> 4)) + (uint64_t)(2 * (ks & 3));
                umma_bf16(tmem + kAttnOCol, a_desc, b_desc, idesc, ks != 0 ? 1u : 0u);
            }
            umma_commit_a(bar_o);
        }
        mbar_wait_a(bar_o, ph);
        tc_fence_after();
        const float inv = 1.f / sum;
        const bool store = q0 + row < T;
        __nv_bfloat16* dst = out + ((size_t)(row0 + q0 + row)) * C + (size_t)h * HD;
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 32) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(lane_base + kAttnOCol + (uint32_t)c0, ov);
            tmem_ld_wait();
            if (store) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 pk;
                    pk.x = pack_bf16x2(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
                    pk.y = pack_bf16x2(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
                    pk.z = pack_bf16x2(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
                    pk.w = pack_bf16x2(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + c0 + i) = pk;
                }
            }
        }
        tc_fence_before();
        __syncthreads();                                  // S / O / P / Q are reused by the next query tile
        tc_fence_after();
    }
    if (warp == 0) tmem_dealloc(tmem, 256);
}
inline size_t attn_tc_smem(int Tp) {
    const size_t kreg = ((size_t)Tp * 128 + 1023) & ~(size_t)1023;
    const size_t atoms = (Tp + 63) / 64;
    return 16384 + 2 * kreg + atoms * (8192 + 16384) + 64;
}

}  // namespace

// The dynamic shared-memory limit is a property of the (process-global) kernel, not of a handle: it is raised to the
// device maximum once, so handles with different token counts can coexist in one process in any creation order.
cudaError_t attention_configure(int T) {
    if (attn_smem(T) > 227 * 1024) return cudaErrorInvalidValue;
    static cudaError_t once = [] {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }();
    return once;
}

cudaError_t launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int N, int T, int H, int C, float scale,
                             cudaStream_t st) {
    if (C != H * HD) return cudaErrorInvalidValue;
    const int Tp = attn_tp(T);
    TMAE_CARVEOUT_ONCE(attention_kernel);
    return launch_k(attention_kernel, dim3(N * H), dim3(kAttnThreads), attn_smem(T), st, true, qkv, out, T, Tp, H, C,
                    scale * 1.4426950408889634f);
}

// tcgen05 path: supported for head_dim 64 and T <= 192 (S row + O fit 256 TMEM columns); the tensor maps (Q: box 64 x 128,
// K: 64 x Tp, both SWIZZLE_128B; V: 64 x Tp unswizzled) are built by the plan.
// MEASURED (B200, batch 64): correct (tests/test_gpu_engine.py::test_attention_matches_torch) but SLOWER than the mma.sync
// kernel - 25 us vs 18 us per launch at T = 65, 106 us vs 43 us at T = 145.  One (image, head) is a strictly serial chain
// (TMA -> S MMA -> softmax -> P to smem -> PV MMA -> store, every hop a barrier round trip) and its 84-128 KB of shared
// memory + 256 TMEM columns allow 1-2 CTAs per SM, against 6 CTAs / 30 warps of the mma.sync kernel: at these sizes
// (0.8-3 % of the path's flops) occupancy beats the faster pipe.  It is therefore opt-in (TMAE_TC_ATTN=1); making it win
// needs a warp-specialised kernel that pipelines several heads per CTA.
bool attention_tc_supported(int T) { return attn_tp(T) <= kAttnTcMaxTp; }
bool attention_tc_eligible(int T) {
    static const bool on = getenv("TMAE_TC_ATTN") != nullptr;
    return on && attention_tc_supported(T);
}
int attention_tc_tp(int T) { return attn_tp(T); }
cudaError_t launch_attention_tc(const CUtensorMap* map_q, const CUtensorMap* map_k, const CUtensorMap* map_v, __nv_bfloat16* out,
                                int N, int T, int H, int C, float scale, cudaStream_t st) {
    if (C != H * HD) return cudaErrorInvalidValue;
    const int Tp = attn_tp(T);
    TMAE_CARVEOUT_ONCE(attention_tc_kernel);
    return launch_k(attention_tc_kernel, dim3(N * H), dim3(kAttnTcThreads), attn_tc_smem(Tp), st, true, *map_q, *map_k, *map_v, out, T,
                    Tp, H, C, scale * 1.4426950408889634f);
}

cudaError_t launch_attention_f32(const __nv_bfloat16* qkv, long long qkv_lo, __nv_bfloat16* out, long long out_lo, int N,
                                 int T, int H, int C, float scale, cudaStream_t st) {
    if (C != H * HD || qkv_lo == 0 || out_lo == 0) return cudaErrorInvalidValue;
    if (attn_f32_smem(T) > 227 * 1024) return cudaErrorInvalidValue;
    TMAE_CARVEOUT_ONCE(attention_f32_kernel);
    return launch_k(attention_f32_kernel, dim3(N * H), dim3(32 * kAttnF32Warps), attn_f32_smem(T), st, true, qkv, qkv_lo, out,
                    out_lo, T, H, C, scale);
}

}  // namespace tmae
