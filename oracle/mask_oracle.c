/* ORACLE (test infrastructure only; nothing under oracle/ is linked into the product library).
 *
 * Torch-free, bit-level CPU restatement of the reference's score-guided patch ordering
 *   MCM.get_ids_shuffle   /root/reference/models/Compression/MCM.py:364-423
 * One call handles a batch [N, L] of fp32 scores and writes the int64 permutation [N, L]
 * (reference return value, MCM.py:423).
 *
 * The reference routine is written with torch CPU ops on fp32 tensors, so "bit exact" means
 * reproducing what those ATen CPU kernels compute (torch 2.11 here).  The non-obvious parts,
 * each pinned by tests/test_mask_oracle.py against the verbatim reference function:
 *   - torch.quantile (linear): rank = q*(n-1) in fp32, lerp = fma(w, hi-lo, lo) for w < 0.5 and
 *     fma(-(1-w)... see lerp_f() below (ATen native/Lerp.h compiled with FMA contraction).
 *   - Tensor.mean(): ATen cascade_sum (native/cpu/SumKernel.cpp): 4 interleaved partial sums of
 *     8-lane vectors (the sum kernel is the AVX2 build even when torch reports AVX512; measured),
 *     scalar tail, then lanes in order; tensors shorter than 8 use 4 interleaved scalar partial
 *     sums.  Then sum / n in fp32.
 *   - F.softmax on 9 values: max, Sleef expf_u10 (FMA build) of (x - max), lane sum, then
 *     multiply by (1 / sum).  The lane-sum order is the only ISA dependent step: sequential for
 *     the AVX512 kernel (9 < 16 lanes, `isa` = 16), lane0 += x[8] then a 4-2-1 shuffle tree for
 *     AVX2 (`isa` = 8).  Both were measured against torch 2.11 (ATEN_CPU_CAPABILITY=avx2 for 8).
 *   - torch.round = round-half-even; NaN -> int32 gives INT_MIN (x86 cvttps2dq), and
 *     `len - num_to_keep` is evaluated in int32 with wrap-around (0-dim int32 tensor arithmetic).
 *   - `group_score[int(start):]` follows Python slice semantics (negative start wraps once).
 *   - Counter() keeps first-appearance order; each value expands to its first `freq` indices.
 *
 * dbg (optional, 32 floats / sample): thr[0..9) | means[9..19) | softmax[19..28) | counts[28..32) unused.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NQ 9
#define NG 10

static const uint32_t kPercentileBits[NQ] = {           /* torch.arange(0.1, 0.91, 0.1, float32), MCM.py:381 */
    0x3dcccccdu, 0x3e4ccccdu, 0x3e99999au, 0x3ecccccdu, 0x3f000000u,
    0x3f19999au, 0x3f333333u, 0x3f4ccccdu, 0x3f666666u};

static float bits2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }

static int cmp_float(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* ATen native/Lerp.h: weight < 0.5 ? self + weight*diff : end - diff*(1-weight); both arms contract to one FMA. */
static float lerp_f(float lo, float hi, float w) {
    float diff = hi - lo;
    if (fabsf(w) < 0.5f) return fmaf(w, diff, lo);
    return fmaf(-diff, 1.0f - w, hi);
}

/* ---- ATen cascade_sum restatement (native/cpu/SumKernel.cpp: multi_row_sum / row_sum / vectorized_inner_sum) ---- */
#define MAXW 16
static int ceil_log2_i64(int64_t x) { int r = 0; while (((int64_t)1 << r) < x) ++r; return r; }

/* nrows = 4 interleaved rows, each row accumulates W-lane vectors.  data: element (i, row k, lane l) at
 * data[(i*4 + k) * W + l]; size = number of i.  out[4][W]. */
static void multi_row_sum4(const float* data, int64_t size, int W, float out[4][MAXW]) {
    enum { num_levels = 4 };
    int lp = ceil_log2_i64(size) / num_levels; if (lp < 4) lp = 4;
    const int64_t level_step = (int64_t)1 << lp, level_mask = level_step - 1;
    static float acc[num_levels][4][MAXW];
    memset(acc, 0, sizeof(acc));
    int64_t i = 0;
    for (; i + level_step <= size;) {
        for (int64_t j = 0; j < level_step; ++j, ++i)
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < W; ++l) acc[0][k][l] += data[(i * 4 + k) * W + l];
        for (int j = 1; j < num_levels; ++j) {
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < W; ++l) { acc[j][k][l] += acc[j - 1][k][l]; acc[j - 1][k][l] = 0.0f; }
            const int64_t mask = level_mask << (j * lp);
            if ((i & mask) != 0) break;
        }
    }
    for (; i < size; ++i)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < W; ++l) acc[0][k][l] += data[(i * 4 + k) * W + l];
    for (int j = 1; j < num_levels; ++j)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < W; ++l) acc[0][k][l] += acc[j][k][l];
    for (int k = 0; k < 4; ++k)
        for (int l = 0; l < W; ++l) out[k][l] = acc[0][k][l];
}

/* row_sum over `size` W-lane vectors -> one W-lane vector */
static void row_sum(const float* data, int64_t size, int W, float out[MAXW]) {
    const int64_t size_ilp = size / 4;
    float ps[4][MAXW];
    multi_row_sum4(data, size_ilp, W, ps);
    for (int64_t i = size_ilp * 4; i < size; ++i)
        for (int l = 0; l < W; ++l) ps[0][l] += data[i * W + l];
    for (int k = 1; k < 4; ++k)
        for (int l = 0; l < W; ++l) ps[0][l] += ps[k][l];
    for (int l = 0; l < W; ++l) out[l] = ps[0][l];
}

static float aten_sum_f32(const float* x, int64_t n, int W) {
    if (n >= W) {                                   /* vectorized_inner_sum */
        const int64_t vec_size = n / W;
        float vacc[MAXW];
        row_sum(x, vec_size, W, vacc);
        float fin = 0.0f;
        for (int64_t k = vec_size * W; k < n; ++k) fin += x[k];
        for (int l = 0; l < W; ++l) fin += vacc[l];
        return 0.0f + fin;
    }
    float s[MAXW];                                  /* scalar row_sum (W = 1) */
    row_sum(x, n, 1, s);
    return 0.0f + s[0];
}

/* ---- exp variants ---- */
/* Sleef_expf_u10 (FMA build) == ATen Vectorized<float>::exp, used by the CPU softmax kernel */
static float exp_sleef_u10(float d) {
    const float R_LN2f = 1.442695040888963407359924681001892137426645954152985934135449406931f;
    const float L2Uf = 0.693145751953125f, L2Lf = 1.428606765330187045e-06f;
    float qf = rintf(d * R_LN2f);
    int q = (int)qf;
    float s = fmaf(qf, -L2Uf, d);
    s = fmaf(qf, -L2Lf, s);
    float u = 0.000198527617612853646278381f;
    u = fmaf(u, s, 0.00139304355252534151077271f);
    u = fmaf(u, s, 0.00833336077630519866943359f);
    u = fmaf(u, s, 0.0416664853692054748535156f);
    u = fmaf(u, s, 0.166666671633720397949219f);
    u = fmaf(u, s, 0.5f);
    u = 1.0f + fmaf(s * s, u, s);
    /* vldexp2: u * 2^(q>>1) * 2^(q - (q>>1)) */
    int q1 = q >> 1, q2 = q - q1;
    u = u * bits2f((uint32_t)(q1 + 127) << 23) * bits2f((uint32_t)(q2 + 127) << 23);
    if (d < -104.0f) u = 0.0f;
    if (d > 100.0f) u = INFINITY;
    return u;
}

static void softmax9(const float* x, int W, float* out) {
    /* max: NaN propagates (vec::maximum) */
    float m = x[0]; int has_nan = isnan(x[0]);
    for (int i = 1; i < NQ; ++i) { if (isnan(x[i])) has_nan = 1; if (x[i] > m) m = x[i]; }
    if (has_nan) { for (int i = 0; i < NQ; ++i) out[i] = NAN; return; }
    float e[NQ];
    for (int i = 0; i < NQ; ++i) e[i] = exp_sleef_u10(x[i] - m);
    float sum;
    if (W == 16) {                       /* size < Vec::size(): vec_reduce_all(fun, vec, size): sequential into lane 0 */
        sum = e[0];
        for (int i = 1; i < NQ; ++i) sum = sum + e[i];
    } else {                             /* AVX2: lane0 += e[8]; then 128-bit / 64-bit / 32-bit shuffle tree */
        float l[8];
        for (int i = 0; i < 8; ++i) l[i] = e[i];
        l[0] = l[0] + e[8];
        float a0 = l[0] + l[4], a1 = l[1] + l[5], a2 = l[2] + l[6], a3 = l[3] + l[7];
        float b0 = a0 + a2, b1 = a1 + a3;
        sum = b0 + b1;
    }
    const float rcp = 1.0f / sum;
    for (int i = 0; i < NQ; ++i) out[i] = e[i] * rcp;
}

static int32_t float_to_int32_x86(float r) {            /* cvttps2dq: NaN / out-of-range -> 0x80000000 */
    if (isnan(r) || r >= 2147483648.0f || r < -2147483648.0f) return INT32_MIN;
    return (int32_t)r;
}

int tmae_oracle_ids_shuffle(const float* scores, int N, int L, int K, int isa, int64_t* out, float* dbg) {
    if (K > L || L <= 0 || (isa != 8 && isa != 16)) return 1;
    const int W = 8;                                   /* lanes of ATen's sum kernel (see header) */
    float* sorted = (float*)malloc(sizeof(float) * L);
    float* uniq = (float*)malloc(sizeof(float) * L);
    float* grp = (float*)malloc(sizeof(float) * L);
    float* vals = (float*)malloc(sizeof(float) * (size_t)L * 2);
    int* cat = (int*)malloc(sizeof(int) * L);
    char* chosen = (char*)malloc(L);
    char* seen = (char*)malloc((size_t)L * 2);
    if (!sorted || !uniq || !grp || !vals || !cat || !chosen || !seen) return 2;

    for (int n = 0; n < N; ++n) {
        const float* sc = scores + (size_t)n * L;
        int64_t* ord = out + (size_t)n * L;
        float* d = dbg ? dbg + (size_t)n * 32 : NULL;

        /* unique() : sorted distinct values (MCM.py:384) */
        memcpy(sorted, sc, sizeof(float) * L);
        qsort(sorted, L, sizeof(float), cmp_float);
        int nu = 0;
        for (int i = 0; i < L; ++i) if (i == 0 || sorted[i] != sorted[i - 1]) uniq[nu++] = sorted[i];

        /* quantile(linear) (MCM.py:383-384) */
        float thr[NQ];
        for (int i = 0; i < NQ; ++i) {
            float rank = bits2f(kPercentileBits[i]) * (float)(nu - 1);
            int64_t below = (int64_t)rank;
            float w = rank - (float)below;
            int64_t above = (int64_t)ceilf(rank);
            thr[i] = lerp_f(uniq[below], uniq[above], w);
        }
        /* bucketize, right=False: index of first thr >= v  (MCM.py:387) */
        int gsize[NG] = {0};
        for (int i = 0; i < L; ++i) {
            int c = 0;
            while (c < NQ && thr[c] < sc[i]) ++c;
            cat[i] = c; gsize[c]++;
        }
        /* group means (MCM.py:390-393) */
        float means[NG];
        for (int g = 0; g < NG; ++g) {
            int m = 0;
            for (int i = 0; i < L; ++i) if (cat[i] == g) grp[m++] = sc[i];
            means[g] = aten_sum_f32(grp, m, W) / (float)m;          /* 0/0 = NaN for an empty group */
        }
        float sm[NQ];
        softmax9(means, isa, sm);                                      /* MCM.py:399-400 */
        const int n_top = gsize[9];
        const int new_target = K - n_top;                             /* :401 */
        int32_t cnt[NQ];
        for (int g = 0; g < NQ; ++g) cnt[g] = float_to_int32_x86(nearbyintf(sm[g] * (float)new_target));  /* :402 */
        if (d) {
            for (int i = 0; i < NQ; ++i) d[i] = thr[i];
            for (int i = 0; i < NG; ++i) d[9 + i] = means[i];
            for (int i = 0; i < NQ; ++i) d[19 + i] = sm[i];
        }

        /* keep_values list (MCM.py:396, 405-408) */
        int nv = 0;
        for (int i = 0; i < L; ++i) if (cat[i] == 9) vals[nv++] = sc[i];
        for (int g = 0; g < NQ; ++g) {
            int m = 0;
            for (int i = 0; i < L; ++i) if (cat[i] == g) grp[m++] = sc[i];
            qsort(grp, m, sizeof(float), cmp_float);
            int32_t start = (int32_t)((uint32_t)m - (uint32_t)cnt[g]);      /* int32 wrap-around */
            int64_t s = start;
            if (s < 0) { s += m; if (s < 0) s = 0; }                        /* Python slice */
            if (s > m) s = m;
            for (int64_t i = s; i < m; ++i) vals[nv++] = grp[i];
        }
        /* Counter -> indices (MCM.py:410-416) */
        memset(chosen, 0, L);
        memset(seen, 0, nv);
        int no = 0;
        for (int p = 0; p < nv; ++p) {
            if (seen[p]) continue;
            float v = vals[p];
            int freq = 0;
            for (int r = p; r < nv; ++r) if (vals[r] == v) { seen[r] = 1; ++freq; }
            for (int i = 0; i < L && freq > 0; ++i)
                if (sc[i] == v) { ord[no++] = i; chosen[i] = 1; --freq; }
        }
        /* remaining indices ascending (MCM.py:418-420) */
        for (int i = 0; i < L; ++i) if (!chosen[i]) ord[no++] = i;
        if (no != L) return 3;
    }
    free(sorted); free(uniq); free(grp); free(vals); free(cat); free(chosen); free(seen);
    return 0;
}

/* Exported for unit tests of the ATen-kernel restatements. */
float tmae_oracle_sum_f32(const float* x, int n, int W) { return aten_sum_f32(x, n, W); }
void tmae_oracle_softmax9(const float* x, int W, float* out) { softmax9(x, W, out); }
float tmae_oracle_exp(float x) { return exp_sleef_u10(x); }
