/*
 * ORACLE - TEST INFRASTRUCTURE ONLY. Never linked or loaded by the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * C restatement of the reference's exact (brute-force) CPU search, i.e. what
 * `usearch.index.Index.search(..., exact=True)` does for the B1/Hamming and NPHD metrics behind
 *   /root/reference/iscc_search/indexes/usearch/index.py:2037          (ShardedNphdIndex.search)
 *   /root/reference/iscc_search/indexes/simprint/usearch_core.py:165   (ShardedIndex128.search)
 * The native code lives in the un-vendored `usearch-iscc==2.24.6` (uv.lock:2491-2499), so this
 * follows the published formula (docs/explanation/similarity-search.md:24-32):
 *     NPHD(a,b) = popcount(a[:m]^b[:m]) / (8*m),  m = min(len a, len b) bytes
 * Algorithm restated: linear scan of zero-padded fixed-stride rows (+1 length byte per row),
 * one bounded top-k buffer per query, threads over queries.
 * Order: (h/n ascending as exact rational, key ascending) - the tie rule is this project's
 * definition (PARITY UNPINNED for ties, see oracle/nphd_oracle.py header).
 *
 * Validated against oracle/nphd_oracle.py (numpy) in tests/test_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXB 32

typedef struct {
    uint64_t dnum;   /* h * (L / n) with L = common multiple -> exact integer order key */
    uint64_t khi, klo;
    uint32_t row;
    uint16_t h, n;
} cand_t;

static inline int cand_less(const cand_t* a, const cand_t* b) {
    if (a->dnum != b->dnum) return a->dnum < b->dnum;
    if (a->khi != b->khi) return a->khi < b->khi;
    return a->klo < b->klo;
}

/* max-heap on (dnum, key): root = current worst of the best k */
static void heap_sift_down(cand_t* hp, size_t n, size_t i) {
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && cand_less(&hp[m], &hp[l])) m = l;
        if (r < n && cand_less(&hp[m], &hp[r])) m = r;
        if (m == i) return;
        cand_t t = hp[i]; hp[i] = hp[m]; hp[m] = t;
        i = m;
    }
}
static void heap_sift_up(cand_t* hp, size_t i) {
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (!cand_less(&hp[p], &hp[i])) return;
        cand_t t = hp[i]; hp[i] = hp[p]; hp[p] = t;
        i = p;
    }
}
static int cand_cmp_qsort(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (cand_less(x, y)) return -1;
    if (cand_less(y, x)) return 1;
    return 0;
}

/* 8*lcm(1..32): every h/(8m) scaled by this is an exact integer < 2^63 (h <= 256) */
static const uint64_t LCM_BITS = 8ull * 144403552893600ull;

static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* Hamming distance over the first m bytes (1..32) of two 32-byte zero padded rows: four masked 64-bit popcounts,
 * no branches (the prefix mask of every length is tabulated once). */
static uint64_t g_mask[MAXB + 1][4];
static int g_mask_ready = 0;
static void init_masks(void) {
    for (uint32_t m = 0; m <= MAXB; m++)
        for (uint32_t w = 0; w < 4; w++) {
            uint32_t lo = 8 * w;
            uint64_t mask = 0;
            if (m >= lo + 8) mask = ~0ull;
            else if (m > lo) mask = (1ull << (8 * (m - lo))) - 1; /* little-endian load: first bytes are the low ones */
            g_mask[m][w] = mask;
        }
    g_mask_ready = 1;
}
static inline uint32_t prefix_hamming(const uint8_t* a, const uint64_t q[4], uint32_t m) {
    const uint64_t* mk = g_mask[m];
    return (uint32_t)(__builtin_popcountll((load64(a) ^ q[0]) & mk[0]) + __builtin_popcountll((load64(a + 8) ^ q[1]) & mk[1]) +
                      __builtin_popcountll((load64(a + 16) ^ q[2]) & mk[2]) + __builtin_popcountll((load64(a + 24) ^ q[3]) & mk[3]));
}

/*
 * codes:   n rows x 32 bytes, zero padded;  lens: n bytes (1..32)
 * keys_hi: n (uint64 keys, or high half of big-endian 128-bit keys); keys_lo: NULL or n
 * queries: q rows x 32 bytes; qlens: q
 * thr_num/thr_den: keep only h/n <= thr_num/thr_den (thr_den == 0 -> no threshold)
 * outputs (q x k, row major): row index into the store (int64, -1 padded), h, nbits; counts[q]
 * returns 0, or -1 on bad arguments
 */
int oracle_nphd_topk(const uint8_t* codes, const uint8_t* lens, const uint64_t* keys_hi, const uint64_t* keys_lo,
                     size_t n, const uint8_t* queries, const uint8_t* qlens, size_t q, uint32_t k,
                     uint32_t thr_num, uint32_t thr_den, int64_t* out_rows, uint16_t* out_h, uint16_t* out_n,
                     uint32_t* counts, int n_threads) {
    if (k < 1) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    int bad = 0;
    if (!g_mask_ready) init_masks();
#pragma omp parallel for schedule(dynamic, 1)
    for (long qi = 0; qi < (long)q; qi++) {
        cand_t* hp = (cand_t*)malloc(sizeof(cand_t) * k);
        size_t cnt = 0;
        const uint8_t* qv = queries + (size_t)qi * MAXB;
        uint32_t ql = qlens[qi];
        if (ql < 1 || ql > MAXB || !hp) { bad = 1; counts[qi] = 0; free(hp); continue; }
        uint64_t scale[MAXB + 1];
        for (uint32_t m = 1; m <= MAXB; m++) scale[m] = LCM_BITS / (8ull * m);
        uint64_t qw[4];
        for (int w = 0; w < 4; w++) qw[w] = load64(qv + 8 * w);
        uint64_t worst = UINT64_MAX; /* dnum of heap root once the heap is full */
        for (size_t i = 0; i < n; i++) {
            uint32_t m = lens[i] < ql ? lens[i] : ql;
            uint32_t h = prefix_hamming(codes + i * MAXB, qw, m);
            uint64_t dnum = (uint64_t)h * scale[m];
            if (dnum > worst) continue;                       /* cheap reject once k are held */
            if (thr_den && (uint64_t)h * thr_den > (uint64_t)thr_num * 8ull * m) continue;
            cand_t c; c.dnum = dnum; c.khi = keys_hi[i]; c.klo = keys_lo ? keys_lo[i] : 0; c.row = (uint32_t)i;
            c.h = (uint16_t)h; c.n = (uint16_t)(8 * m);
            if (cnt < k) {
                hp[cnt] = c; heap_sift_up(hp, cnt); cnt++;
                if (cnt == k) worst = hp[0].dnum;
            } else if (cand_less(&c, &hp[0])) {
                hp[0] = c; heap_sift_down(hp, cnt, 0); worst = hp[0].dnum;
            }
        }
        qsort(hp, cnt, sizeof(cand_t), cand_cmp_qsort);
        for (size_t j = 0; j < k; j++) {
            size_t o = (size_t)qi * k + j;
            if (j < cnt) { out_rows[o] = hp[j].row; out_h[o] = hp[j].h; out_n[o] = hp[j].n; }
            else { out_rows[o] = -1; out_h[o] = 0; out_n[o] = 0; }
        }
        counts[qi] = (uint32_t)cnt;
        free(hp);
    }
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic rows of SURVEY.md 8d, restating iscc_search_b200/synth.py (the definition; tests/test_oracle.py
 * checks this C copy against it): row i, 64-bit word w = splitmix64(seed ^ (4i + w)); length = lengths[mix(i) % n];
 * key = bijective mix of i (key_mode 0) or the simprint chunk pointer asset8 | offset4 | size4 of chunk i % cpa of
 * asset i / cpa (key_mode 1, lmdb_ops.py:30-49). dup_every > 0: row i with i % dup_every == dup_every - 1 repeats
 * the code of row i - dup_back (shared chunks -> exact duplicates exist for the equality join).
 * Lets the checker cover stores that do not fit host memory (1 B rows): rows are regenerated block by block.
 */
static inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct {
    uint64_t seed;
    const uint8_t* lengths;
    uint32_t n_lengths, key_mode, cpa, dup_every, dup_back;
} synth_t;

static inline void synth_row(const synth_t* g, uint64_t i, uint8_t* code /*32*/, uint8_t* len, uint64_t* khi, uint64_t* klo) {
    uint32_t L = g->lengths[g->n_lengths > 1 ? splitmix64(i ^ ((g->seed + 0x1234567ull))) % g->n_lengths : 0];
    uint64_t src = (g->dup_every && i % g->dup_every == g->dup_every - 1 && i >= g->dup_back) ? i - g->dup_back : i;
    uint64_t w[4];
    for (uint32_t k = 0; k < 4; k++) w[k] = splitmix64(g->seed ^ (src * 4 + k));
    memcpy(code, w, 32);
    memset(code + L, 0, 32 - L);
    *len = (uint8_t)L;
    const uint64_t kc = (g->seed * 0x51ED27ull + 0xA5A5A5A5ull);
    if (g->key_mode == 0) { *khi = splitmix64(i ^ kc); *klo = 0; }
    else { *khi = splitmix64((i / g->cpa) ^ kc); *klo = ((uint64_t)((i % g->cpa) * 4096u) << 32) | 4096u; }
}

int oracle_synth_rows(uint64_t seed, uint64_t start, size_t n, const uint8_t* lengths, uint32_t n_lengths, uint32_t key_mode,
                      uint32_t cpa, uint32_t dup_every, uint32_t dup_back, uint64_t* keys_hi, uint64_t* keys_lo, uint8_t* codes,
                      uint8_t* lens, int n_threads) {
    if (!lengths || n_lengths < 1 || (key_mode == 1 && cpa < 1)) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    synth_t g = {seed, lengths, n_lengths, key_mode, cpa, dup_every, dup_back};
#pragma omp parallel for schedule(static)
    for (long j = 0; j < (long)n; j++) {
        uint64_t hi, lo;
        synth_row(&g, start + (uint64_t)j, codes + (size_t)j * MAXB, lens + j, &hi, &lo);
        keys_hi[j] = hi;
        if (keys_lo) keys_lo[j] = lo;
    }
    return 0;
}

/* Exact top-k over n_rows synthetic rows WITHOUT materialising them: threads take blocks of rows, regenerate them,
 * score every query against the block and keep one bounded heap per (thread, query); the heaps are merged at the end.
 * Same order and threshold rules as oracle_nphd_topk. Outputs: keys (hi, lo), h, nbits, counts. */
int oracle_synth_topk(uint64_t seed, uint64_t n_rows, const uint8_t* lengths, uint32_t n_lengths, uint32_t key_mode, uint32_t cpa,
                      uint32_t dup_every, uint32_t dup_back, const uint8_t* queries, const uint8_t* qlens, size_t q, uint32_t k,
                      uint32_t thr_num, uint32_t thr_den, uint64_t* out_khi, uint64_t* out_klo, uint16_t* out_h, uint16_t* out_n,
                      uint32_t* counts, int n_threads) {
    if (k < 1 || !lengths || n_lengths < 1 || (key_mode == 1 && cpa < 1)) return -1;
    for (size_t i = 0; i < q; i++) if (qlens[i] < 1 || qlens[i] > MAXB) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    const int T = omp_get_max_threads();
#else
    const int T = 1;
#endif
    if (!g_mask_ready) init_masks();
    synth_t g = {seed, lengths, n_lengths, key_mode, cpa, dup_every, dup_back};
    enum { BLK = 2048 };
    cand_t* heaps = (cand_t*)malloc(sizeof(cand_t) * (size_t)T * q * k);
    size_t* cnts = (size_t*)calloc((size_t)T * q, sizeof(size_t));
    uint64_t* worst = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)T * q);
    if (!heaps || !cnts || !worst) { free(heaps); free(cnts); free(worst); return -1; }
    for (size_t i = 0; i < (size_t)T * q; i++) worst[i] = UINT64_MAX;
    uint64_t scale[MAXB + 1];
    scale[0] = 0;
    for (uint32_t m = 1; m <= MAXB; m++) scale[m] = LCM_BITS / (8ull * m);
    const long n_blocks = (long)((n_rows + BLK - 1) / BLK);
#pragma omp parallel
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        uint8_t* codes = (uint8_t*)malloc((size_t)BLK * MAXB);
        uint8_t lens[BLK];
        uint64_t khi[BLK], klo[BLK];
#pragma omp for schedule(dynamic, 4)
        for (long b = 0; b < n_blocks; b++) {
            const uint64_t r0 = (uint64_t)b * BLK;
            const uint32_t bn = (uint32_t)((n_rows - r0 < BLK) ? n_rows - r0 : BLK);
            for (uint32_t j = 0; j < bn; j++) synth_row(&g, r0 + j, codes + (size_t)j * MAXB, lens + j, khi + j, klo + j);
            for (size_t qi = 0; qi < q; qi++) {
                const uint32_t ql = qlens[qi];
                uint64_t qw[4];
                for (int w = 0; w < 4; w++) qw[w] = load64(queries + qi * MAXB + 8 * w);
                cand_t* hp = heaps + ((size_t)t * q + qi) * k;
                size_t cnt = cnts[(size_t)t * q + qi];
                uint64_t wst = worst[(size_t)t * q + qi];
                for (uint32_t j = 0; j < bn; j++) {
                    const uint32_t m = lens[j] < ql ? lens[j] : ql;
                    const uint32_t h = prefix_hamming(codes + (size_t)j * MAXB, qw, m);
                    const uint64_t dnum = (uint64_t)h * scale[m];
                    if (dnum > wst) continue;
                    if (thr_den && (uint64_t)h * thr_den > (uint64_t)thr_num * 8ull * m) continue;
                    cand_t c; c.dnum = dnum; c.khi = khi[j]; c.klo = klo[j]; c.row = 0; c.h = (uint16_t)h; c.n = (uint16_t)(8 * m);
                    if (cnt < k) {
                        hp[cnt] = c; heap_sift_up(hp, cnt); cnt++;
                        if (cnt == k) wst = hp[0].dnum;
                    } else if (cand_less(&c, &hp[0])) {
                        hp[0] = c; heap_sift_down(hp, cnt, 0); wst = hp[0].dnum;
                    }
                }
                cnts[(size_t)t * q + qi] = cnt;
                worst[(size_t)t * q + qi] = wst;
            }
        }
        free(codes);
    }
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (long qi = 0; qi < (long)q; qi++) {
        size_t total = 0;
        for (int t = 0; t < T; t++) total += cnts[(size_t)t * q + qi];
        cand_t* all = (cand_t*)malloc(sizeof(cand_t) * (total ? total : 1));
        if (!all) { bad = 1; counts[qi] = 0; continue; }
        size_t o = 0;
        for (int t = 0; t < T; t++) {
            memcpy(all + o, heaps + ((size_t)t * q + qi) * k, sizeof(cand_t) * cnts[(size_t)t * q + qi]);
            o += cnts[(size_t)t * q + qi];
        }
        qsort(all, total, sizeof(cand_t), cand_cmp_qsort);
        const size_t c = total < k ? total : k;
        for (size_t j = 0; j < k; j++) {
            const size_t oo = (size_t)qi * k + j;
            if (j < c) { out_khi[oo] = all[j].khi; if (out_klo) out_klo[oo] = all[j].klo; out_h[oo] = all[j].h; out_n[oo] = all[j].n; }
            else { out_khi[oo] = ~0ull; if (out_klo) out_klo[oo] = ~0ull; out_h[oo] = 0; out_n[oo] = 0; }
        }
        counts[qi] = (uint32_t)c;
        free(all);
    }
    free(heaps); free(cnts); free(worst);
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------------
 * The CPU arm bench.py times beside the GPU: the same exact scan, laid out the way a tuned
 * CPU implementation would be - length-bucketed structure of arrays of 64-bit word planes, so a query reads only
 * min(Lq, Lb) bytes of every row, scored 8 rows at a time with AVX-512 VPOPCNTDQ (runtime dispatch, scalar popcount
 * otherwise), one bounded top-k heap per query, OpenMP over queries. Results are identical to oracle_nphd_topk
 * (tests/test_oracle.py), which stays the plain restatement the parity tests rely on.
 */
#include <immintrin.h>

typedef struct {
    uint32_t len;        /* code length of the bucket, bytes */
    uint32_t words;      /* ceil(len / 8) */
    size_t n, cap;       /* rows, rows rounded up to 8 */
    uint64_t* planes;    /* [words][cap], zero padded */
    uint64_t* khi; uint64_t* klo; uint32_t* row;   /* [n] */
} soa_bucket_t;

typedef struct { soa_bucket_t b[MAXB + 1]; int has_lo; } soa_t;

void oracle_soa_free(void* h) {
    soa_t* s = (soa_t*)h;
    if (!s) return;
    for (int L = 1; L <= MAXB; L++) { free(s->b[L].planes); free(s->b[L].khi); free(s->b[L].klo); free(s->b[L].row); }
    free(s);
}

void* oracle_soa_build(const uint8_t* codes, const uint8_t* lens, const uint64_t* keys_hi, const uint64_t* keys_lo, size_t n) {
    if (n > 0xffffffffull) return NULL;
    soa_t* s = (soa_t*)calloc(1, sizeof(soa_t));
    if (!s) return NULL;
    s->has_lo = keys_lo != NULL;
    size_t cnt[MAXB + 1] = {0};
    for (size_t i = 0; i < n; i++) { if (lens[i] < 1 || lens[i] > MAXB) { free(s); return NULL; } cnt[lens[i]]++; }
    for (int L = 1; L <= MAXB; L++) {
        soa_bucket_t* b = &s->b[L];
        b->len = (uint32_t)L; b->words = (uint32_t)((L + 7) / 8); b->n = 0; b->cap = (cnt[L] + 7) / 8 * 8;
        if (!cnt[L]) continue;
        b->planes = (uint64_t*)aligned_alloc(64, sizeof(uint64_t) * b->words * b->cap);
        b->khi = (uint64_t*)malloc(sizeof(uint64_t) * cnt[L]);
        b->klo = keys_lo ? (uint64_t*)malloc(sizeof(uint64_t) * cnt[L]) : NULL;
        b->row = (uint32_t*)malloc(sizeof(uint32_t) * cnt[L]);
        if (!b->planes || !b->khi || !b->row || (keys_lo && !b->klo)) { oracle_soa_free(s); return NULL; }
        memset(b->planes, 0, sizeof(uint64_t) * b->words * b->cap);
    }
    for (size_t i = 0; i < n; i++) {
        soa_bucket_t* b = &s->b[lens[i]];
        const size_t r = b->n++;
        uint8_t tmp[MAXB] = {0};
        memcpy(tmp, codes + i * MAXB, lens[i]);
        for (uint32_t w = 0; w < b->words; w++) b->planes[(size_t)w * b->cap + r] = load64(tmp + 8 * w);
        b->khi[r] = keys_hi[i];
        if (keys_lo) b->klo[r] = keys_lo[i];
        b->row[r] = (uint32_t)i;
    }
    return s;
}

/* distances of 8 consecutive rows of a bucket over the first nw words (last word masked) */
__attribute__((target("avx512f,avx512vpopcntdq"))) static inline __m512i soa_dist8_avx512(const soa_bucket_t* b, size_t i, const uint64_t* qw,
                                                                                        uint32_t nw, uint64_t mask_last) {
    __m512i acc = _mm512_setzero_si512();
    for (uint32_t w = 0; w < nw; w++) {
        __m512i x = _mm512_xor_si512(_mm512_load_si512((const void*)(b->planes + (size_t)w * b->cap + i)), _mm512_set1_epi64((long long)qw[w]));
        if (w == nw - 1) x = _mm512_and_si512(x, _mm512_set1_epi64((long long)mask_last));
        acc = _mm512_add_epi64(acc, _mm512_popcnt_epi64(x));
    }
    return acc;
}

typedef struct { cand_t* hp; size_t cnt; uint32_t k; uint64_t worst; } topk_t;

static inline void topk_offer(topk_t* t, const cand_t* c) {
    if (t->cnt < t->k) {
        t->hp[t->cnt] = *c; heap_sift_up(t->hp, t->cnt); t->cnt++;
        if (t->cnt == t->k) t->worst = t->hp[0].dnum;
    } else if (cand_less(c, &t->hp[0])) {
        t->hp[0] = *c; heap_sift_down(t->hp, t->cnt, 0); t->worst = t->hp[0].dnum;
    }
}

__attribute__((target("avx512f,avx512vpopcntdq"))) static void soa_scan_avx512(const soa_bucket_t* b, const uint64_t* qw, uint32_t m, uint64_t scale,
                                                                             uint32_t thr_h, topk_t* t) {
    const uint32_t nw = (m + 7) / 8;
    const uint64_t mask_last = (m & 7) ? ((1ull << (8 * (m & 7))) - 1) : ~0ull;
    uint64_t hb = t->worst == UINT64_MAX ? 8ull * m : t->worst / scale;   /* rows beyond hb cannot enter the heap */
    if (hb > thr_h) hb = thr_h;
    for (size_t i = 0; i < b->n; i += 8) {
        const __m512i d = soa_dist8_avx512(b, i, qw, nw, mask_last);
        __mmask8 pass = _mm512_cmple_epu64_mask(d, _mm512_set1_epi64((long long)hb));
        if (i + 8 > b->n) pass &= (__mmask8)((1u << (b->n - i)) - 1u);
        if (!pass) continue;
        uint64_t dv[8];
        _mm512_storeu_si512((void*)dv, d);
        while (pass) {
            const int j = __builtin_ctz(pass);
            pass &= (__mmask8)(pass - 1);
            const size_t r = i + (size_t)j;
            cand_t c; c.dnum = dv[j] * scale; c.khi = b->khi[r]; c.klo = b->klo ? b->klo[r] : 0; c.row = b->row[r];
            c.h = (uint16_t)dv[j]; c.n = (uint16_t)(8 * m);
            topk_offer(t, &c);
        }
        hb = t->worst == UINT64_MAX ? 8ull * m : t->worst / scale;
        if (hb > thr_h) hb = thr_h;
    }
}

static void soa_scan_scalar(const soa_bucket_t* b, const uint64_t* qw, uint32_t m, uint64_t scale, uint32_t thr_h, topk_t* t) {
    const uint32_t nw = (m + 7) / 8;
    const uint64_t mask_last = (m & 7) ? ((1ull << (8 * (m & 7))) - 1) : ~0ull;
    for (size_t r = 0; r < b->n; r++) {
        uint64_t h = 0;
        for (uint32_t w = 0; w < nw; w++) {
            uint64_t x = b->planes[(size_t)w * b->cap + r] ^ qw[w];
            if (w == nw - 1) x &= mask_last;
            h += (uint64_t)__builtin_popcountll(x);
        }
        if (h > thr_h) continue;
        const uint64_t dnum = h * scale;
        if (dnum > t->worst) continue;
        cand_t c; c.dnum = dnum; c.khi = b->khi[r]; c.klo = b->klo ? b->klo[r] : 0; c.row = b->row[r]; c.h = (uint16_t)h; c.n = (uint16_t)(8 * m);
        topk_offer(t, &c);
    }
}

/* 1 when the AVX-512 VPOPCNTDQ kernel is used on this machine */
int oracle_soa_isa(void) {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vpopcntdq") ? 1 : 0;
}

int oracle_soa_topk(const void* h, const uint8_t* queries, const uint8_t* qlens, size_t q, uint32_t k, uint32_t thr_num, uint32_t thr_den,
                    int64_t* out_rows, uint16_t* out_h, uint16_t* out_n, uint32_t* counts, int n_threads, int force_scalar) {
    const soa_t* s = (const soa_t*)h;
    if (!s || k < 1) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    const int fast = !force_scalar && oracle_soa_isa();
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (long qi = 0; qi < (long)q; qi++) {
        const uint32_t ql = qlens[qi];
        topk_t t; t.hp = (cand_t*)malloc(sizeof(cand_t) * k); t.cnt = 0; t.k = k; t.worst = UINT64_MAX;
        if (ql < 1 || ql > MAXB || !t.hp) { bad = 1; counts[qi] = 0; free(t.hp); continue; }
        uint8_t qpad[MAXB] = {0};
        memcpy(qpad, queries + (size_t)qi * MAXB, ql);
        uint64_t qw[4];
        for (int w = 0; w < 4; w++) qw[w] = load64(qpad + 8 * w);
        for (uint32_t L = 1; L <= MAXB; L++) {
            const soa_bucket_t* b = &s->b[L];
            if (!b->n) continue;
            const uint32_t m = L < ql ? L : ql;
            const uint64_t scale = LCM_BITS / (8ull * m);
            const uint32_t thr_h = thr_den ? (uint32_t)(((uint64_t)thr_num * 8ull * m) / thr_den) : 8u * m;   /* h <= thr * n */
            if (fast) soa_scan_avx512(b, qw, m, scale, thr_h, &t);
            else soa_scan_scalar(b, qw, m, scale, thr_h, &t);
        }
        qsort(t.hp, t.cnt, sizeof(cand_t), cand_cmp_qsort);
        for (size_t j = 0; j < k; j++) {
            const size_t o = (size_t)qi * k + j;
            if (j < t.cnt) { out_rows[o] = t.hp[j].row; out_h[o] = t.hp[j].h; out_n[o] = t.hp[j].n; }
            else { out_rows[o] = -1; out_h[o] = 0; out_n[o] = 0; }
        }
        counts[qi] = (uint32_t)t.cnt;
        free(t.hp);
    }
    return bad ? -1 : 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
