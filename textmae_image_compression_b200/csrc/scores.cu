// Patch-score generation on the GPU: the reference's offline generator, generate_scores_file.py:19-31,
// which costs 0.3-1.0 s per image in Python pixel loops:
//   s_map = Division_Merge_Segmented(img)  utils/map.py:6-53   quadtree split (mean / std(ddof=1) / 95 % rule) + threshold merge,
//                                                              crop [1:-1, 1:-1], cv2.resize (bilinear, 8-bit fixed point)
//   t_map = laplacian(img)                 utils/map.py:56-60  3x3 Laplacian of the ALREADY SEGMENTED image (the reference's
//                                                              first call works in place), |.| saturated to 8 bits, cv2.resize
//   score = cal_patch_score(t) * cal_patch_score(s), min-max normalised   utils/distribution.py:5-16, generate_scores_file.py:24-31
// Integer / byte work, HBM- and latency-bound (0.4 MB per image): four small kernels, every image of the batch in one launch.
//   1. score_judge_large_kernel  nodes of the top quadtree levels (> 2048 pixels): several CTAs per node build per-warp
//                                shared-memory histograms, merge them through a global histogram; the last CTA decides
//   2. score_judge_small_kernel  one warp per node for the deeper levels
//      (the split decisions of ALL levels are evaluated on the original pixels: regions are disjoint, so a node's decision
//       never depends on another node's merge - the recursion is only needed to know which decisions are *used*)
//   3. score_segment_kernel      per pixel: walk the decision tree from the root, threshold at the leaf
//   4. score_patch_kernel        one CTA per 16x16 output patch: both resizes evaluated on the fly (the Laplacian inside the
//                                t_map taps), block sums -> int(mean) product; the last CTA of an image normalises.
// Decisions are taken in IEEE float64 like numpy's (mean = S / n exact-rounded; the variance sum is accumulated per grey
// level instead of numpy's pairwise pixel order: a few ulps apart, which can only matter when (v - mean) equals 2 std to
// ~1e-15 relative - not observed on Kodak or the synthetic sets; tests/test_gpu_scores.py is bit-exact on all of them).
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace tmae {

namespace {

constexpr int kJudgeThreads = 256;
constexpr int kJudgeWarps = kJudgeThreads / 32;

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {      // fixed butterfly order: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// origin of node (i, j) of level d: every split moves by the child size of that level (utils/map.py:37-40)
__device__ __forceinline__ void node_origin(const ScoreGeom& g, int d, int i, int j, int& oy, int& ox) {
    oy = 0; ox = 0;
    for (int k = 1; k <= d; ++k) {
        if ((i >> (d - k)) & 1) oy += g.h[k];
        if ((j >> (d - k)) & 1) ox += g.w[k];
    }
}

// Division_Judge (utils/map.py:6-23) from the grey-level histogram of the node: cnt[v] pixels of value v, n pixels.
// Called by one warp; lane l owns levels [8 l, 8 l + 8).
__device__ __forceinline__ bool judge_from_hist_warp(const unsigned* hist, int n, int lane) {
    unsigned c[8];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k] = hist[lane * 8 + k]; s += (unsigned long long)c[k] * (unsigned)(lane * 8 + k); }
    s = warp_sum_u64(s);
    const double mean = __ddiv_rn((double)s, (double)n);                    // np.mean: exact integer sum / n
    double sq = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double dev = __dsub_rn((double)(lane * 8 + k), mean);
        sq = __dadd_rn(sq, __dmul_rn((double)c[k], __dmul_rn(dev, dev)));
    }
    sq = warp_sum_f64(sq);
    const double sd2 = __dmul_rn(2.0, sqrt(__ddiv_rn(sq, (double)(n - 1))));   // 2 * np.std(ddof=1)
    unsigned op = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double dev = __dsub_rn((double)(lane * 8 + k), mean);
        if (dev < sd2) op += c[k];
    }
    op = warp_sum_u32(op);
    return __ddiv_rn((double)op, (double)n) >= 0.95;
}

__global__ void __launch_bounds__(kJudgeThreads)
score_judge_large_kernel(const uint8_t* __restrict__ gray, const __grid_constant__ ScoreGeom g, unsigned* __restrict__ ghist, unsigned* __restrict__ tickets,
                         uint8_t* __restrict__ flags) {
    __shared__ unsigned hist[kJudgeWarps][256];
    __shared__ unsigned is_last;
    const int img = blockIdx.y;
    int d = 0, rem = blockIdx.x;
    while (rem >= g.large_ctas[d]) { rem -= g.large_ctas[d]; ++d; }         // level of this CTA
    const int chunks = g.chunks[d];
    const int node = rem / chunks, chunk = rem - node * chunks;
    const int i = node >> d, j = node & ((1 << d) - 1);
    int oy, ox;
    node_origin(g, d, i, j, oy, ox);
    const int h = g.h[d], w = g.w[d];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int k = tid; k < kJudgeWarps * 256; k += kJudgeThreads) (&hist[0][0])[k] = 0;
    __syncthreads();
    const int rows_per = (h + chunks - 1) / chunks;
    const int r0 = chunk * rows_per, r1 = min(h, r0 + rows_per);
    const uint8_t* base = gray + (size_t)img * g.H * g.W + (size_t)oy * g.W + ox;
    // one warp per row (coalesced byte loads along the row), rows dealt round-robin to the warps: no per-pixel division
    for (int r = r0 + warp; r < r1; r += kJudgeWarps) {
        const uint8_t* rowp = base + (size_t)r * g.W;
        for (int c = tid & 31; c < w; c += 32) atomicAdd(&hist[warp][rowp[c]], 1u);
    }
    __syncthreads();
    unsigned cnt = 0;
#pragma unroll
    for (int k = 0; k < kJudgeWarps; ++k) cnt += hist[k][tid];
    const int gnode = g.large_node_off[d] + node;
    unsigned* gh = ghist + ((size_t)img * g.large_nodes + gnode) * 256;
    if (chunks > 1) {
        if (cnt) atomicAdd(&gh[tid], cnt);
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = (atomicAdd(&tickets[(size_t)img * g.large_nodes + gnode], 1u) == (unsigned)(chunks - 1));
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        cnt = __ldcg(&gh[tid]);
    }
    hist[0][tid] = cnt;
    __syncthreads();
    if (warp == 0) {
        const bool uniform = judge_from_hist_warp(hist[0], h * w, tid);
        if (tid == 0) flags[(size_t)img * g.flag_total + g.flag_off[d] + node] = uniform ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kJudgeThreads)
score_judge_small_kernel(const uint8_t* __restrict__ gray, const __grid_constant__ ScoreGeom g, uint8_t* __restrict__ flags) {
    __shared__ unsigned hist[kJudgeWarps][256];
    const int img = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int rem = blockIdx.x * kJudgeWarps + warp;                              // node index over the small levels
    int d = g.large_levels;
    while (d < g.levels && rem >= (1 << (2 * d))) { rem -= 1 << (2 * d); ++d; }
    if (d >= g.levels) return;                                              // whole warp
    const int node = rem, i = node >> d, j = node & ((1 << d) - 1);
    int oy, ox;
    node_origin(g, d, i, j, oy, ox);
    const int h = g.h[d], w = g.w[d];
#pragma unroll
    for (int k = 0; k < 8; ++k) hist[warp][lane * 8 + k] = 0;
    __syncwarp();
    const uint8_t* base = gray + (size_t)img * g.H * g.W + (size_t)oy * g.W + ox;
    {   // lane -> pixel e = lane, lane + 32, ...: (r, c) advanced incrementally instead of divided per pixel
        int r = lane / w, c = lane - r * w;
        const int dr = 32 / w, dc = 32 - dr * w;
        for (int e = lane; e < h * w; e += 32) {
            atomicAdd(&hist[warp][base[(size_t)r * g.W + c]], 1u);
            r += dr; c += dc;
            if (c >= w) { c -= w; ++r; }
        }
    }
    __syncwarp();
    const bool uniform = judge_from_hist_warp(hist[warp], h * w, lane);
    if (lane == 0) flags[(size_t)img * g.flag_total + g.flag_off[d] + node] = uniform ? 1 : 0;
}

// Recursion + Merge (utils/map.py:27-42) for one pixel: descend while the node was split; a pixel on the odd last row /
// column of a split node belongs to no child and keeps its value.
__device__ __forceinline__ uint8_t segment_pixel(const ScoreGeom& g, const uint8_t* __restrict__ fl, int y, int x, uint8_t v) {
    int d = 0, node_i = 0, node_j = 0, oy = 0, ox = 0;
    while (d < g.levels && !fl[g.flag_off[d] + (node_i << d) + node_j]) {
        const int nh = g.h[d + 1], nw = g.w[d + 1];
        const int cy = (y - oy) >= nh ? 1 : 0, cx = (x - ox) >= nw ? 1 : 0;
        oy += cy * nh; ox += cx * nw;
        if (y - oy >= nh || x - ox >= nw) return v;              // odd last row / column of a split node: never visited
        node_i = node_i * 2 + cy; node_j = node_j * 2 + cx;
        ++d;
    }
    return (v > 60 && v < 150) ? 0 : 255;                         // Merge (utils/map.py:27-31)
}
// VEC = 4: W % 4 == 0, a thread handles four consecutive pixels of one row with 32-bit loads / stores
template <int VEC>
__global__ void __launch_bounds__(256)
score_segment_kernel(const uint8_t* __restrict__ gray, const __grid_constant__ ScoreGeom g, const uint8_t* __restrict__ flags, uint8_t* __restrict__ seg) {
    const int img = blockIdx.y;
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (p >= g.H * g.W) return;
    const int y = p / g.W, x = p - y * g.W;
    const uint8_t* fl = flags + (size_t)img * g.flag_total;
    const size_t off = (size_t)img * g.H * g.W + p;
    if (VEC == 4) {
        const uchar4 v = *reinterpret_cast<const uchar4*>(gray + off);
        uchar4 o;
        o.x = segment_pixel(g, fl, y, x, v.x); o.y = segment_pixel(g, fl, y, x + 1, v.y);
        o.z = segment_pixel(g, fl, y, x + 2, v.z); o.w = segment_pixel(g, fl, y, x + 3, v.w);
        *reinterpret_cast<uchar4*>(seg + off) = o;
    } else {
        seg[off] = segment_pixel(g, fl, y, x, gray[off]);
    }
}

// cv2.resize (INTER_LINEAR, 8-bit) tap of one output coordinate: source indices and 11-bit coefficients.
__device__ __forceinline__ void resize_tap(int dcoord, double scale, int ssize, bool zero_edges, int& s0, int& s1, int& a0, int& a1) {
    float f = (float)__dsub_rn(__dmul_rn((double)dcoord + 0.5, scale), 0.5);
    int s = (int)floorf(f);
    f = f - (float)s;
    if (zero_edges) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    }
    a0 = __float2int_rn((1.f - f) * 2048.f);
    a1 = __float2int_rn(f * 2048.f);
    s0 = min(max(s, 0), ssize - 1);
    s1 = min(max(s + 1, 0), ssize - 1);
}
__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}
__device__ __forceinline__ int lap_abs(const uint8_t* __restrict__ seg, int H, int W, int y, int x) {
    const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H), xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
    const int v = 2 * ((int)seg[(size_t)ym * W + xm] + (int)seg[(size_t)ym * W + xp] + (int)seg[(size_t)yp * W + xm] +
                       (int)seg[(size_t)yp * W + xp]) - 8 * (int)seg[(size_t)y * W + x];
    return min(abs(v), 255);
}
__device__ __forceinline__ int resize_vertical(int r0, int r1, int b0, int b1) {
    const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    return min(max(v, 0), 255);
}

__global__ void __launch_bounds__(256)
score_patch_kernel(const uint8_t* __restrict__ seg_all, const __grid_constant__ ScoreGeom g, int* __restrict__ prod, unsigned* __restrict__ tickets,
                   float* __restrict__ scores, uint8_t* __restrict__ s_map, uint8_t* __restrict__ t_map) {
    __shared__ int red_s[8], red_t[8];
    __shared__ unsigned is_last;
    __shared__ int red_mn[8], red_mx[8];
    const int img = blockIdx.y, patch = blockIdx.x;
    const int side = g.S / 16, L = side * side;
    const int pr = patch / side, pc = patch - pr * side;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Y = pr * 16 + (tid >> 4), X = pc * 16 + (tid & 15);
    const uint8_t* seg = seg_all + (size_t)img * g.H * g.W;
    int sx0, sx1, ax0, ax1, sy0, sy1, ay0, ay1;
    // s_map: resize of seg[1:-1, 1:-1]
    resize_tap(X, g.scale_x_crop, g.W - 2, true, sx0, sx1, ax0, ax1);
    resize_tap(Y, g.scale_y_crop, g.H - 2, false, sy0, sy1, ay0, ay1);
    const uint8_t* c0 = seg + (size_t)(sy0 + 1) * g.W + 1;
    const uint8_t* c1 = seg + (size_t)(sy1 + 1) * g.W + 1;
    const int sv = resize_vertical((int)c0[sx0] * ax0 + (int)c0[sx1] * ax1, (int)c1[sx0] * ax0 + (int)c1[sx1] * ax1, ay0, ay1);
    // t_map: resize of |Laplacian(seg)|
    resize_tap(X, g.scale_x_full, g.W, true, sx0, sx1, ax0, ax1);
    resize_tap(Y, g.scale_y_full, g.H, false, sy0, sy1, ay0, ay1);
    const int l00 = lap_abs(seg, g.H, g.W, sy0, sx0), l01 = lap_abs(seg, g.H, g.W, sy0, sx1);
    const int l10 = lap_abs(seg, g.H, g.W, sy1, sx0), l11 = lap_abs(seg, g.H, g.W, sy1, sx1);
    const int tv = resize_vertical(l00 * ax0 + l01 * ax1, l10 * ax0 + l11 * ax1, ay0, ay1);
    if (s_map) s_map[((size_t)img * g.S + Y) * g.S + X] = (uint8_t)sv;
    if (t_map) t_map[((size_t)img * g.S + Y) * g.S + X] = (uint8_t)tv;
    // int(mean) of the 16 x 16 block (utils/distribution.py:10-12), product (generate_scores_file.py:26)
    int ss = sv, ts = tv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); ts += __shfl_xor_sync(0xffffffffu, ts, o); }
    if (lane == 0) { red_s[warp] = ss; red_t[warp] = ts; }
    __syncthreads();
    if (tid == 0) {
        int a = 0, b = 0;
        for (int k = 0; k < 8; ++k) { a += red_s[k]; b += red_t[k]; }
        prod[(size_t)img * L + patch] = (a >> 8) * (b >> 8);
        __threadfence();
        is_last = (atomicAdd(&tickets[img], 1u) == (unsigned)(L - 1));
    }
    __syncthreads();
    if (!is_last) return;
    // the last patch of this image: (total - min) / (max - min) in float64, cast to fp32 (generate_scores_file.py:28-31)
    __threadfence();
    const int* pv = prod + (size_t)img * L;
    int mn = INT_MAX, mx = INT_MIN;
    for (int k = tid; k < L; k += 256) { const int v = __ldcg(&pv[k]); mn = min(mn, v); mx = max(mx, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (lane == 0) { red_mn[warp] = mn; red_mx[warp] = mx; }
    __syncthreads();
    mn = red_mn[0]; mx = red_mx[0];
    for (int k = 1; k < 8; ++k) { mn = min(mn, red_mn[k]); mx = max(mx, red_mx[k]); }
    const double den = (double)(mx - mn);
    for (int k = tid; k < L; k += 256) scores[(size_t)img * L + k] = (float)__ddiv_rn((double)(__ldcg(&pv[k]) - mn), den);
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------------------
bool score_geometry(int H, int W, int S, ScoreGeom* out) {
    if (H < 8 || W < 8 || S < 16 || S % 16 != 0) return false;
    ScoreGeom g = {};
    g.H = H; g.W = W; g.S = S;
    g.h[0] = H; g.w[0] = W;
    int d = 0, off = 0;
    while (d < kScoreMaxLevels && (g.h[d] < g.w[d] ? g.h[d] : g.w[d]) > 5) {    // Recursion splits only while min(h, w) > 5
        g.flag_off[d] = off;
        off += 1 << (2 * d);
        g.h[d + 1] = g.h[d] / 2; g.w[d + 1] = g.w[d] / 2;
        ++d;
    }
    if (d >= kScoreMaxLevels) return false;
    g.levels = d;
    g.flag_total = off;
    g.large_levels = 0;
    g.large_nodes = 0;
    g.large_cta_total = 0;
    for (int k = 0; k < g.levels; ++k) {
        const long long px = (long long)g.h[k] * g.w[k];
        if (px <= 2048 || k != g.large_levels) break;
        int chunks = (int)((px + 16383) / 16384);
        if (chunks > g.h[k]) chunks = g.h[k];
        g.chunks[k] = chunks;
        g.large_node_off[k] = g.large_nodes;
        g.large_ctas[k] = (1 << (2 * k)) * chunks;
        g.large_nodes += 1 << (2 * k);
        g.large_cta_total += g.large_ctas[k];
        g.large_levels = k + 1;
    }
    g.small_nodes = 0;
    for (int k = g.large_levels; k < g.levels; ++k) g.small_nodes += 1 << (2 * k);
    g.scale_x_crop = 1.0 / ((double)S / (double)(W - 2));      // hal::resize: scale = 1. / inv_scale, inv_scale = dsize / ssize
    g.scale_y_crop = 1.0 / ((double)S / (double)(H - 2));
    g.scale_x_full = 1.0 / ((double)S / (double)W);
    g.scale_y_full = 1.0 / ((double)S / (double)H);
    *out = g;
    return true;
}

// workspace layout per call: [zeroed: ghist u32 [n][large_nodes][256] | node tickets u32 [n][large_nodes] | image tickets u32 [n]]
//                            | prod i32 [n][L] | flags u8 [n][flag_total] | segmented u8 [n][H][W]
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
size_t score_workspace_bytes(const ScoreGeom& g, int n) {
    const size_t L = (size_t)(g.S / 16) * (g.S / 16);
    const size_t zeroed = align256(((size_t)n * g.large_nodes * 257 + (size_t)n) * 4);
    return zeroed + align256((size_t)n * L * 4) + align256((size_t)n * g.flag_total) + align256((size_t)n * g.H * g.W);
}

cudaError_t launch_generate_scores(const uint8_t* gray, int n, const ScoreGeom& g, float* scores, uint8_t* s_map, uint8_t* t_map,
                                   uint8_t* seg_out, void* workspace, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t L = (size_t)(g.S / 16) * (g.S / 16);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    const size_t zeroed = align256(((size_t)n * g.large_nodes * 257 + (size_t)n) * 4);
    unsigned* ghist = reinterpret_cast<unsigned*>(ws);
    unsigned* node_tickets = ghist + (size_t)n * g.large_nodes * 256;
    unsigned* img_tickets = node_tickets + (size_t)n * g.large_nodes;
    int* prod = reinterpret_cast<int*>(ws + zeroed);
    uint8_t* flags = ws + zeroed + align256((size_t)n * L * 4);
    uint8_t* seg = flags + align256((size_t)n * g.flag_total);
    cudaError_t e = cudaMemsetAsync(ws, 0, zeroed, st);
    if (e != cudaSuccess) return e;
    if (g.large_cta_total > 0) {
        score_judge_large_kernel<<<dim3(g.large_cta_total, n), kJudgeThreads, 0, st>>>(gray, g, ghist, node_tickets, flags);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (g.small_nodes > 0) {
        score_judge_small_kernel<<<dim3((g.small_nodes + kJudgeWarps - 1) / kJudgeWarps, n), kJudgeThreads, 0, st>>>(gray, g, flags);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    const bool vec4 = g.W % 4 == 0 && (reinterpret_cast<uintptr_t>(gray) & 3) == 0;          // seg is 256-byte aligned in the workspace
    if (vec4) score_segment_kernel<4><<<dim3((g.H * g.W / 4 + 255) / 256, n), 256, 0, st>>>(gray, g, flags, seg);
    else score_segment_kernel<1><<<dim3((g.H * g.W + 255) / 256, n), 256, 0, st>>>(gray, g, flags, seg);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    score_patch_kernel<<<dim3((unsigned)L, n), 256, 0, st>>>(seg, g, prod, img_tickets, scores, s_map, t_map);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (seg_out) e = cudaMemcpyAsync(seg_out, seg, (size_t)n * g.H * g.W, cudaMemcpyDeviceToDevice, st);
    return e;
}

}  // namespace tmae
